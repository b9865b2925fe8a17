"""NumPy float64 restatement of the reference's Jacobian-determinant map -- TEST ORACLE ONLY.

PINNED: follows /root/reference ``eval_reg_with_jacobian.py:62-78`` operation by operation
(4th-order central differences on the interior, ``J[n, c, d] = d u_c / d x_d``, ``+ I``,
``np.linalg.det``, strictly-negative count) and is checked in ``tests/test_oracle_jacobian.py``
against ``tests/golden/jacobian_*.npz``, which were produced by executing those lines of the
reference file itself (``tests/golden/make_jacobian_golden.py``).
"""
import numpy as np


def jacobian_determinant(ddf):
    """ddf: (H, W, D, 1, 3) displacement field (any float dtype; promoted like get_fdata()).

    Returns (det [(H-4)*(W-4)*(D-4)] float64, n_negative int)  -- eval_reg_with_jacobian.py:62-78.
    """
    ddf = np.asarray(ddf, dtype=np.float64)          # nib get_fdata() -> float64 (:51)
    height, width, depth, time_dim, num_channel = ddf.shape
    num_voxel = (height - 4) * (width - 4) * (depth - 4)
    c = ddf[2:-2, 2:-2, 2:-2]
    # :66-68  (u[-2] - 8 u[-1] + 8 u[+1] - u[+2]) / 12 along each axis, evaluated left to right
    dx = ((ddf[:-4, 2:-2, 2:-2] - 8 * ddf[1:-3, 2:-2, 2:-2] + 8 * ddf[3:-1, 2:-2, 2:-2]
           - ddf[4:, 2:-2, 2:-2]) / 12.0).reshape(num_voxel, num_channel)
    dy = ((ddf[2:-2, :-4, 2:-2] - 8 * ddf[2:-2, 1:-3, 2:-2] + 8 * ddf[2:-2, 3:-1, 2:-2]
           - ddf[2:-2, 4:, 2:-2]) / 12.0).reshape(num_voxel, num_channel)
    dz = ((ddf[2:-2, 2:-2, :-4] - 8 * ddf[2:-2, 2:-2, 1:-3] + 8 * ddf[2:-2, 2:-2, 3:-1]
           - ddf[2:-2, 2:-2, 4:]) / 12.0).reshape(num_voxel, num_channel)
    del c
    J = np.stack([dx, dy, dz], 2)                     # :69
    J[:, 0, 0] += 1                                   # :71-73
    J[:, 1, 1] += 1
    J[:, 2, 2] += 1
    det = np.linalg.det(J)                            # :74
    n_negative = int(np.count_nonzero(np.where(det > 0, 0, det)))   # :76-77  (det == 0 not counted)
    return det, n_negative


def summary(det, n_negative):
    """The CSV statistics of eval_reg_with_jacobian.py:78,84-91 (same keys, same order)."""
    res = dict()
    res['percentage_negative_detJa'] = 100 * n_negative / len(det)
    res['median_detJa'] = np.median(det)
    res['mean_detJa'] = np.mean(det)
    res['std_detJa'] = np.std(det)
    res['n_total_detJa'] = len(det)
    res['n_negatives_detJa'] = n_negative
    return res
