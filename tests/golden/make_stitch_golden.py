"""Generate tests/golden/stitch_*.npz by EXECUTING the reference's own stitching function.

Run in the build container only (needs /root/reference):
    python tests/golden/make_stitch_golden.py

Reads lines 214-259 of /root/reference/3d_reg.py (`get_def_field_from_subvol`) at run time and
`exec`s the definition (it needs NumPy only); the tile placement comes from lines 165-207 of the same
file, executed the same way on stand-in volumes.  No reference source is copied into the repository.
"""
import os
import textwrap

import numpy as np

REF = '/root/reference/3d_reg.py'
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_stitch():
    with open(REF) as f:
        lines = f.readlines()
    src = ''.join(lines[213:259])                      # file lines 214..259
    assert src.startswith('def get_def_field_from_subvol') and 'return warp_field' in src
    env = {'np': np}
    exec(compile(src, REF + ':214-259', 'exec'), env)
    return env['get_def_field_from_subvol']


def reference_tiles(shape_in_vol, subvol_size, min_perc_overlap):
    with open(REF) as f:
        lines = f.readlines()
    src = textwrap.dedent(''.join(lines[158:207]))     # file lines 159..207 (body of `if use_subvol:`)
    assert 'nb_sub_x_axis' in src and 'lst_coords_subvol.append' in src
    vol = np.zeros(shape_in_vol, np.float32)
    env = {'np': np, 'model_inference_specs': {'subvol_size': list(subvol_size), 'min_perc_overlap': min_perc_overlap},
           'fx_img_res111': vol, 'mov_img_res111': vol}
    exec(compile(src, REF + ':159-207', 'exec'), env)
    return env['in_shape'], env['lst_coords_subvol']


def make_warps(seed, in_shape, n):
    rng = np.random.default_rng(int(seed))
    return [(rng.standard_normal(tuple(int(d) for d in in_shape) + (3,)) * 2).astype(np.float32) for _ in range(n)]


def main():
    stitch = reference_stitch()
    cases = {'two_by_two': ((24, 20, 28), (16, 16, 16), 0.1), 'coincident_tiles': ((16, 16, 16), (16, 16, 16), 0.1),
             'many_small': ((30, 18, 20), (16, 16, 16), 0.25)}
    for i, (name, (vol_shape, subvol, perc)) in enumerate(cases.items()):
        in_shape, coords = reference_tiles(vol_shape, subvol, perc)
        warps = make_warps(515 + i, in_shape, len(coords))
        out = stitch(in_shape, vol_shape, coords, warps)
        # the tile fields are regenerated from the seed by the tests (keeps the fixture small)
        np.savez_compressed(os.path.join(HERE, 'stitch_%s.npz' % name), vol_shape=np.array(vol_shape),
                            in_shape=np.array(in_shape), subvol=np.array(subvol), perc=np.float64(perc),
                            coords=np.array(coords), seed=np.int64(515 + i), out=out)
        print(name, vol_shape, in_shape, len(coords), 'tiles', out.shape, out.dtype)


if __name__ == '__main__':
    main()
