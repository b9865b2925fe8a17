"""NIfTI-1 codec round trips (the on-disk format at the edge of the hot path)."""
import gzip
import os
import struct

import numpy as np
import pytest

from multimodal_registration_b200 import _nifti
from multimodal_registration_b200.voxelmorph.py import utils as pyutils


@pytest.mark.parametrize('dtype', [np.float32, np.float64, np.uint8, np.int16, np.int32])
@pytest.mark.parametrize('ext', ['.nii', '.nii.gz'])
def test_round_trip(tmp_path, dtype, ext):
    rng = np.random.default_rng(0)
    a = (rng.random((5, 6, 7)) * 100).astype(dtype)
    aff = np.array([[0, -1.5, 0, 10], [2.0, 0, 0, -20], [0, 0, 0.8, 5], [0, 0, 0, 1]])
    p = str(tmp_path / ('v' + ext))
    _nifti.save_nifti(a, p, aff)
    b, aff2, hdr = _nifti.load_nifti(p, return_header=True)
    assert b.dtype == np.dtype(dtype) and b.shape == a.shape
    np.testing.assert_array_equal(a, b)
    np.testing.assert_allclose(aff, aff2, atol=1e-6)
    assert hdr['intent_code'] == 0 and hdr['sform_code'] > 0


def test_vector_field_5d_and_intent(tmp_path):
    f = np.random.default_rng(1).standard_normal((4, 5, 6, 1, 3)).astype(np.float32)
    p = str(tmp_path / 'warp.nii.gz')
    _nifti.save_nifti(f, p, np.eye(4), intent_code=1007)
    g, aff, hdr = _nifti.load_nifti(p, return_header=True)
    np.testing.assert_array_equal(f, g)
    assert hdr['intent_code'] == 1007 and hdr['dim'][0] == 5
    # raw layout check: Fortran order, little endian, data at offset 352
    raw = gzip.open(p, 'rb').read()
    assert struct.unpack('<i', raw[:4])[0] == 348 and raw[344:347] == b'n+1'
    np.testing.assert_array_equal(np.frombuffer(raw, '<f4', offset=352, count=4), f.reshape(-1, order='F')[:4])


def test_load_save_volfile_semantics(tmp_path):
    a = np.random.default_rng(2).random((4, 5, 6, 1)).astype(np.float32)
    p = str(tmp_path / 'im.nii.gz')
    pyutils.save_volfile(a, p, np.diag([1, 2, 3, 1.0]))
    v = pyutils.load_volfile(p, add_batch_axis=True, add_feat_axis=True)
    assert v.shape == (1, 4, 5, 6, 1)                       # squeeze, then batch + feature axes
    v2, aff = pyutils.load_volfile(p, ret_affine=True)
    assert v2.shape == (4, 5, 6)
    np.testing.assert_allclose(aff, np.diag([1, 2, 3, 1.0]))


def test_aff2axcodes():
    assert _nifti.aff2axcodes(np.eye(4)) == ('R', 'A', 'S')
    assert _nifti.aff2axcodes(np.diag([-1, -1, 1, 1.0])) == ('L', 'P', 'S')
    aff = np.array([[0, 0, 1, 0], [-1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1.0]])
    assert _nifti.aff2axcodes(aff) == ('P', 'S', 'R')
