"""Generate tests/golden/sct_perm.npz by EXECUTING the reference's own RAI conversion lines.

Run in the build container only (needs /root/reference):
    python tests/golden/make_sct_golden.py

Reads lines 399-417 of /root/reference/3d_reg.py at run time and `exec`s them with a stand-in `nib` whose
`aff2axcodes` returns the orientation under test (nibabel is not installed; the lines only use its result),
for all 48 axis-code orientations.  No reference source is copied into the repository.
"""
import itertools
import os
import textwrap
import types

import numpy as np

REF = '/root/reference/3d_reg.py'
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    with open(REF) as f:
        lines = f.readlines()
    src = textwrap.dedent(''.join(lines[398:417]))            # file lines 399..417
    assert src.lstrip().startswith('orientation_conv = "RAI"') and 'warp_data_exp[..., 2]' in src
    rng = np.random.default_rng(77)
    warp = rng.standard_normal((3, 4, 5, 3)).astype(np.float32)
    pairs = (('L', 'R'), ('P', 'A'), ('I', 'S'))
    codes, perms, invs, outs = [], [], [], []
    for order in itertools.permutations(range(3)):
        for signs in itertools.product((0, 1), repeat=3):
            ax = tuple(pairs[order[i]][signs[i]] for i in range(3))
            nib = types.SimpleNamespace(aff2axcodes=lambda a, ax=ax: ax)
            env = {'np': np, 'nib': nib, 'fixed_nii': types.SimpleNamespace(affine=np.eye(4)), 'warp_data': warp.copy()}
            exec(compile(src, REF + ':399-417', 'exec'), env)
            codes.append(''.join(ax)); perms.append(env['perm']); invs.append(env['inversion']); outs.append(env['warp_data_exp'])
    np.savez_compressed(os.path.join(HERE, 'sct_perm.npz'), codes=np.array(codes), perm=np.array(perms),
                        inversion=np.array(invs), warp=warp, out=np.stack(outs))
    print(len(codes), 'orientations', np.stack(outs).shape)


if __name__ == '__main__':
    main()
