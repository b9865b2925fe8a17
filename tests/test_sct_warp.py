"""SCT warp convention: oracle and host logic pinned to the reference's own lines
(tests/golden/sct_perm.npz, made by executing 3d_reg.py:399-417); GPU export against the oracle."""
import os

import numpy as np
import pytest

import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import _nifti, sct_warp
from oracle import interp_oracle as io
from oracle import sct_oracle

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'sct_perm.npz'))


def test_oracle_and_host_permutation_match_reference_lines():
    assert len(GOLD['codes']) == 48
    for code, perm, inv, out in zip(GOLD['codes'], GOLD['perm'], GOLD['inversion'], GOLD['out']):
        ax = tuple(str(code))
        assert sct_oracle.rai_permutation(ax) == (list(perm), list(inv))
        assert sct_warp.rai_permutation(ax) == (list(perm), list(inv))
        np.testing.assert_array_equal(sct_oracle.apply(GOLD['warp'], ax), out)


def _affine_for(code, spacing=(1.0, 1.0, 1.0)):
    """An affine whose NEGATION has axis codes `code` (the reference calls aff2axcodes(-affine))."""
    rows = {'R': (0, 1), 'L': (0, -1), 'A': (1, 1), 'P': (1, -1), 'S': (2, 1), 'I': (2, -1)}
    a = np.zeros((4, 4))
    for col, ch in enumerate(code):
        r, s = rows[ch]
        a[r, col] = -s * spacing[col]
    a[3, 3] = 1.0
    return a


def test_affine_helper_roundtrip():
    for code in ('RAS', 'LPI', 'PIR', 'SLA', 'ARI'):
        assert ''.join(_nifti.aff2axcodes(-_affine_for(code))) == code


@pytest.mark.gpu
@pytest.mark.parametrize('code', ['RAS', 'LPI', 'PIR', 'SLA', 'AIL', 'IRP'])
@pytest.mark.parametrize('scale', [1, 2])
def test_to_sct_warp_matches_oracle(code, scale, tmp_path):
    rng = np.random.default_rng(5)
    half = (rng.standard_normal((1, 6, 8, 8, 3)) * 2).astype(np.float32)
    aff = _affine_for(code, (1.0, 0.8, 1.2))
    want = sct_oracle.apply(io.rescale_dense_transform(half, scale)[0], tuple(code))
    got = sct_warp.to_sct_warp(half[0], aff, scale)
    assert got.shape == want.shape and got.dtype == np.float32
    if mrb._lib.exact_order():
        np.testing.assert_array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-4)
    path = str(tmp_path / 'warp.nii.gz')
    sct_warp.save_sct_warp(path, half[0], aff, scale)
    data, affine, hdr = _nifti.load_nifti(path, return_header=True)
    assert hdr['intent_code'] == 1007                      # 3d_reg.py:418
    np.testing.assert_array_equal(np.asarray(data, dtype=np.float32), got)
    np.testing.assert_allclose(affine, aff, atol=1e-6)
