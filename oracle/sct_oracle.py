"""CPU restatement of the SCT warp convention (3d_reg.py:399-417) -- TEST INFRASTRUCTURE ONLY.

Pinned: tests/golden/sct_perm.npz holds the permutation / inversion and a transformed array for every
one of the 48 orientations, produced by executing those reference lines
(tests/golden/make_sct_golden.py)."""
import numpy as np


def rai_permutation(axcodes):
    """3d_reg.py:399-411 with `fx_im_orientation = list(axcodes)`."""
    conv = 'RAI'
    opposite = {'L': 'R', 'R': 'L', 'A': 'P', 'P': 'A', 'I': 'S', 'S': 'I'}
    perm, inversion = [0, 1, 2], [1, 1, 1]
    for i, ch in enumerate(conv):
        try:
            perm[i] = list(axcodes).index(ch)
        except ValueError:
            perm[i] = list(axcodes).index(opposite[ch])
            inversion[i] = -1
    return perm, inversion


def apply(warp_xyz3, axcodes):
    """3d_reg.py:413-417: time axis, then permuted / sign-flipped components."""
    perm, inversion = rai_permutation(axcodes)
    w = np.expand_dims(np.asarray(warp_xyz3), axis=3)
    out = np.copy(w)
    for i in range(3):
        out[..., i] = inversion[i] * w[..., perm[i]]
    return out
