"""Drop-in scripts end to end on the GPU (NIfTI in, NIfTI/CSV out) against the oracle."""
import csv
import os
import runpy
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scripts'))

from multimodal_registration_b200 import _nifti     # noqa: E402
from oracle import interp_oracle as io               # noqa: E402
from oracle import jacobian_oracle as jo             # noqa: E402

pytestmark = pytest.mark.gpu


def _load(name):
    return runpy.run_path(os.path.join(ROOT, 'scripts', name))


def test_eval_reg_with_jacobian_script(tmp_path):
    rng = np.random.default_rng(5)
    f = rng.standard_normal((14, 12, 16, 1, 3)).astype(np.float32)
    aff = np.diag([1.0, 1.0, 2.0, 1.0])
    fp = str(tmp_path / 'sub-01_proc_field_to_T2w.nii.gz')
    _nifti.save_nifti(f, fp, aff, intent_code=1007)
    mod = _load('eval_reg_with_jacobian.py')
    out_csv, out_im = str(tmp_path / 'jac.csv'), str(tmp_path / 'detJa.nii.gz')
    for _ in range(2):                                       # second call appends
        assert mod['main'](['--def-field-path', fp, '--sub-id', 'sub-01', '--out-file', out_csv,
                            '--out-im-path', out_im]) == 0
    det, n_neg = jo.jacobian_determinant(f)
    want = jo.summary(det, n_neg)
    rows = list(csv.reader(open(out_csv)))
    assert rows[0] == ['Timestamp', 'Subject', 'Percentage_negative_detJa[%]', 'Median_detJa', 'Mean_detJa',
                       'Std_detJa', 'N_total_voxels', 'N_voxels_negatives_detJa']
    assert len(rows) == 3 and rows[1][1] == 'sub-01'
    got = [float(v) for v in rows[1][2:]]
    exp = [want[k] for k in ('percentage_negative_detJa', 'median_detJa', 'mean_detJa', 'std_detJa', 'n_total_detJa',
                             'n_negatives_detJa')]
    np.testing.assert_allclose(got, exp, rtol=1e-5, atol=1e-5)
    assert got[4] == exp[4] and got[5] == exp[5]
    vol, aff2 = _nifti.load_nifti(out_im)
    assert vol.shape == (10, 8, 12, 1) and vol.dtype == np.float64
    np.testing.assert_allclose(vol.reshape(-1), det, rtol=0, atol=1e-4)
    np.testing.assert_allclose(aff2, aff)


@pytest.mark.parametrize('interp', ['linear', 'nearest'])
def test_gen_apply_def_field_script(tmp_path, interp):
    rng = np.random.default_rng(6)
    vol = (rng.random((16, 20, 24)) * 100).astype(np.float32)
    ip = str(tmp_path / 'im.nii.gz')
    _nifti.save_nifti(vol, ip, np.eye(4))
    mod = _load('gen_apply_def_field.py')
    res = str(tmp_path / 'res')
    assert mod['main'](['--im-path', ip, '--res-dir', res, '--def-scales', '4', '8', '--def-max-std', '2',
                        '--interp', interp, '--seed', '3']) == 0
    field, _ = _nifti.load_nifti(os.path.join(res, 'deformation_field.nii.gz'))
    moved, _ = _nifti.load_nifti(os.path.join(res, 'moved_im.nii.gz'))
    assert field.shape == (16, 20, 24, 3) and moved.shape == vol.shape
    assert 0.05 < np.abs(field).mean() < 6            # a non-trivial smooth field
    want = io.transform(vol[..., None], field.astype(np.float32), interp)[..., 0]
    if interp == 'nearest':
        np.testing.assert_array_equal(moved, want)
    else:
        np.testing.assert_allclose(moved, want, rtol=1e-5, atol=1e-4 * 100)
