"""Pin the stitching oracle to golden vectors produced by the reference's own function
(3d_reg.py:214-259 and the tile placement of :159-207, executed by tests/golden/make_stitch_golden.py)."""
import glob
import importlib.util
import os

import numpy as np
import pytest

from oracle import stitch_oracle as so

HERE = os.path.dirname(__file__)
GOLDEN = sorted(glob.glob(os.path.join(HERE, 'golden', 'stitch_*.npz')))
_spec = importlib.util.spec_from_file_location('make_stitch_golden', os.path.join(HERE, 'golden', 'make_stitch_golden.py'))
_gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_gen)


def load_case(path):
    g = np.load(path)
    coords = [tuple(int(v) for v in c) for c in g['coords']]
    warps = _gen.make_warps(int(g['seed']), g['in_shape'], len(coords))
    return g, coords, warps


def test_goldens_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_function(path):
    g, coords, warps = load_case(path)
    out = so.get_def_field_from_subvol(tuple(g['in_shape']), tuple(g['vol_shape']), coords, warps)
    assert out.dtype == np.float64
    np.testing.assert_array_equal(out, g['out'])                    # same ops, same order: bit-exact
    # tile placement restatement agrees with the reference's loop
    in_shape = tuple(int(np.ceil(s // 16)) * 16 for s in g['subvol'])
    assert so.tile_coords(tuple(g['vol_shape']), in_shape, float(g['perc'])) == coords


def test_weights_and_single_tile_known_answers():
    w = so.pyramid_weights((8, 8, 8))
    assert w.shape == (8, 8, 8) and w[4, 4, 4] == 1.0 and w.min() == 1 - 4 / 5
    f = np.random.default_rng(0).standard_normal((8, 8, 8, 3))
    out = so.get_def_field_from_subvol((8, 8, 8), (8, 8, 8), [(0, 8, 0, 8, 0, 8)], [f])
    np.testing.assert_allclose(out, f, rtol=1e-15)                  # one tile: weights cancel
    out = so.get_def_field_from_subvol((8, 8, 8), (12, 8, 8), [(0, 8, 0, 8, 0, 8)], [f])
    assert np.all(out[8:] == 0)                                     # uncovered voxels stay 0
