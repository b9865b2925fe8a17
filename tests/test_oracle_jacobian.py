"""Pin the Jacobian oracle to golden vectors produced by the reference's own lines
(eval_reg_with_jacobian.py:62-78, executed by tests/golden/make_jacobian_golden.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import jacobian_oracle as jo

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'jacobian_*.npz')))


def test_goldens_present():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_lines(path):
    g = np.load(path)
    det, n_neg = jo.jacobian_determinant(g['field'][:, :, :, None, :])
    assert det.dtype == np.float64
    np.testing.assert_array_equal(det, g['det'])          # same ops, same order: bit-exact
    assert n_neg == int(g['n_neg'])
    s = jo.summary(det, n_neg)
    assert s['percentage_negative_detJa'] == float(g['pct'])
    assert s['median_detJa'] == float(g['median'])
    assert s['mean_detJa'] == float(g['mean'])
    assert s['std_detJa'] == float(g['std'])


def test_affine_field_known_answer():
    # u = A x  ->  det(I + A) everywhere in the interior (4th-order stencil is exact on linear u)
    A = np.array([[0.10, -0.05, 0.02], [0.03, -0.20, 0.07], [-0.04, 0.06, 0.15]])
    X, Y, Z = 9, 8, 10
    g = np.stack(np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing='ij'), -1).astype(np.float64)
    u = g @ A.T
    det, n_neg = jo.jacobian_determinant(u[:, :, :, None, :])
    np.testing.assert_allclose(det, np.linalg.det(np.eye(3) + A), rtol=0, atol=1e-12)
    assert n_neg == 0 and det.size == (X - 4) * (Y - 4) * (Z - 4)


def test_zero_field_and_reflection():
    z = np.zeros((6, 6, 6, 1, 3))
    det, n = jo.jacobian_determinant(z)
    assert np.all(det == 1.0) and n == 0
    # u = -2x along axis 0 -> I + J = diag(-1, 1, 1): every voxel folded
    g = np.arange(7, dtype=np.float64)
    u = np.zeros((7, 6, 6, 1, 3))
    u[..., 0, 0] = -2 * g[:, None, None]
    det, n = jo.jacobian_determinant(u)
    np.testing.assert_allclose(det, -1.0, atol=1e-12)
    assert n == det.size
