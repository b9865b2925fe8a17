"""NumPy float64 restatement of the reference's sub-volume stitching -- TEST ORACLE ONLY.

PINNED: follows /root/reference ``3d_reg.py:214-259`` ``get_def_field_from_subvol`` (duplicated at
``bids_registration.py:226-271`` and ``bids_two_steps_registration.py:226-271``) operation by operation
and is checked in ``tests/test_oracle_stitch.py`` against ``tests/golden/stitch_*.npz``, produced by
executing that function of the reference file itself (``tests/golden/make_stitch_golden.py``).
"""
import numpy as np


def pyramid_weights(model_in_shape):
    """3d_reg.py:221-226: 1 - max(|x|,|y|,|z|) / (max + 1) on the grid [-s/2, s/2)."""
    x, y, z = model_in_shape[0] // 2, model_in_shape[1] // 2, model_in_shape[2] // 2
    grid = np.mgrid[-x:x, -y:y, -z:z]
    w_map = np.maximum(np.abs(grid[0]), np.abs(grid[1]))
    w_map = np.maximum(w_map, np.abs(grid[2]))
    return 1 - w_map / (np.max(w_map) + 1)


def get_def_field_from_subvol(model_in_shape, im_shape, lst_coords_subvol, lst_warp_subvol):
    """Weighted average of the overlapping tile fields (3d_reg.py:228-259), float64 like the reference."""
    w_map = pyramid_weights(model_in_shape)
    sum_weights = np.zeros((im_shape[0], im_shape[1], im_shape[2]))
    for (x0, x1, y0, y1, z0, z1) in lst_coords_subvol:               # :232-238
        sum_weights[x0:x1, y0:y1, z0:z1] += w_map
    sum_weights[sum_weights == 0] = 1                                # :246
    warp_field = np.zeros((im_shape[0], im_shape[1], im_shape[2], 3))
    for (x0, x1, y0, y1, z0, z1), warp in zip(lst_coords_subvol, lst_warp_subvol):
        w_rel = w_map / sum_weights[x0:x1, y0:y1, z0:z1]             # :250-252
        for i in range(3):                                           # :255-258 (tile order, one add per tile)
            warp_field[x0:x1, y0:y1, z0:z1, i] += w_rel * warp[..., i]
    return warp_field


def tile_coords(shape_in_vol, in_shape, min_perc):
    """Tile placement of 3d_reg.py:165-207 (host logic; returns the (min, max) tuples)."""
    if min_perc >= 1:
        min_perc = min_perc / 100 if min_perc / 100 < 1 else 0.1
    elif min_perc <= 0:
        min_perc = 0.1
    nb = [int(shape_in_vol[d] / (in_shape[d] - min_perc * in_shape[d])) + 1 for d in range(3)]
    ov = [0.0, 0.0, 0.0]
    for d in range(3):
        if nb[d] > 1:
            ov[d] = (in_shape[d] - (shape_in_vol[d] / nb[d])) * (nb[d] / (nb[d] - 1))
    coords = []
    x_max = y_max = z_max = 0
    for i in range(nb[0]):
        x_min = 0 if i == 0 else int(x_max - ov[0])
        x_max = int(x_min + in_shape[0])
        for j in range(nb[1]):
            y_min = 0 if j == 0 else int(y_max - ov[1])
            y_max = int(y_min + in_shape[1])
            for k in range(nb[2]):
                z_min = 0 if k == 0 else int(z_max - ov[2])
                z_max = int(z_min + in_shape[2])
                coords.append((x_min, x_max, y_min, y_max, z_min, z_max))
    return coords
