"""Evaluation metrics on the GPU (multimodal_registration_b200.metrics) against
  * the goldens produced by executing the reference's own lines (eval_reg_with_mi.py:16-74,
    eval_reg_on_sc_seg.py:80-124; tests/golden/make_metrics_golden.py), and
  * oracle/metrics_oracle.py (np.histogramdd) on larger volumes, float32 and float64,
and the two drop-in scripts end to end (NIfTI in, CSV out)."""
import csv
import glob
import os
import runpy
import sys

import numpy as np
import pytest
import torch

import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import _nifti, metrics
from oracle import metrics_oracle as mo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'metrics_*.npz')))


def crop(a, box):
    x0, y0, z0, x1, y1, z1 = [int(v) for v in box]
    return np.ascontiguousarray(a[x0:x1 + 1, y0:y1 + 1, z0:z1 + 1])


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_metrics_against_reference_outputs(path):
    g = np.load(path)
    box = metrics.detect_zero_padding(g['moving'])
    np.testing.assert_array_equal(np.array(box), g['box'])
    fx, moving, moved = (crop(g[k], box) for k in ('fx', 'moving', 'moved'))
    hist = metrics.joint_histogram(fx, moved).cpu().numpy()
    np.testing.assert_array_equal(hist, g['hist_fx_moved'].astype(np.int64))          # integer counts: exact
    for key, (a, b) in {'nmi_fx_moving': (fx, moving), 'nmi_fx_moved': (fx, moved), 'nmi_moving_moved': (moving, moved)}.items():
        np.testing.assert_allclose(metrics.normalized_mutual_information(a, b), g[key], rtol=1e-12)
    np.testing.assert_allclose(metrics.normalized_mutual_information(fx, moved, bins=7), g['nmi_bins7'], rtol=1e-12)
    for tag in ('moving', 'moved'):
        m = metrics.overlap_metrics(g['seg_fx'], g['seg_' + tag])
        for ours, ref in (('dice', 'dice'), ('jaccard', 'jacc'), ('sensitivity', 'sens'), ('precision', 'prec'),
                          ('specificity', 'spec'), ('accuracy', 'acc'), ('TP', 'TP'), ('FP', 'FP'), ('TN', 'TN'), ('FN', 'FN')):
            assert m[ours] == g['sc_%s_%s' % (ref, tag)], (ours, tag)              # sums of 0 / 1 values: exact


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('shape', [(64, 48, 80), (160, 160, 192)])
def test_joint_histogram_matches_histogramdd(shape, dtype):
    """Bit-exact bin counts at a BASELINE-size volume, including samples that sit exactly on bin edges."""
    rng = np.random.default_rng(shape[0])
    a = rng.random(shape).astype(dtype)
    b = (0.7 * a + 0.3 * rng.random(shape)).astype(dtype)
    a.ravel()[::97] = np.round(a.ravel()[::97] * 100) / 100            # many values on / next to the edges
    b.ravel()[::89] = b.max()
    # float32 inputs are binned as their float64 values (get_fdata() semantics: the reference only ever sees float64)
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    want = mo.joint_histogram(a64, b64, 100).astype(np.int64)
    got = metrics.joint_histogram(a, b, 100).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    assert got.sum() == a.size
    np.testing.assert_allclose(metrics.normalized_mutual_information(a, b), mo.normalized_mutual_information(a64, b64), rtol=1e-12)
    # device tensors are taken as they are
    got_t = metrics.joint_histogram(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), 100).cpu().numpy()
    np.testing.assert_array_equal(got_t, want)


def test_histogram_of_constant_image_and_integer_labels():
    a = np.full((8, 9, 10), 3.0)                                   # degenerate range: np.histogramdd widens it by 0.5
    b = np.random.default_rng(1).integers(0, 5, a.shape).astype(np.float64)
    np.testing.assert_array_equal(metrics.joint_histogram(a, b, 10).cpu().numpy(), mo.joint_histogram(a, b, 10).astype(np.int64))
    seg = np.random.default_rng(2).integers(0, 3, (20, 20, 20)).astype(np.uint8)      # integer dtypes go through float64
    np.testing.assert_array_equal(metrics.joint_histogram(seg, seg, 3).cpu().numpy(), mo.joint_histogram(seg, seg, 3).astype(np.int64))


def test_overlap_metrics_large_and_empty():
    rng = np.random.default_rng(3)
    fx = (rng.random((96, 96, 64)) > 0.7).astype(np.float64)
    mv = (rng.random((96, 96, 64)) > 0.6).astype(np.float64)
    got, want = metrics.overlap_metrics(fx, mv), mo.overlap_metrics(fx, mv)
    for k in ('dice', 'jaccard', 'sensitivity', 'precision', 'specificity', 'accuracy', 'TP', 'FP', 'TN', 'FN'):
        assert got[k] == want[k], k
    empty = metrics.overlap_metrics(fx, np.zeros_like(fx))         # no foreground in the other map
    assert empty['TP'] == 0 and empty['dice'] == 0 and np.isnan(empty['precision'])


def test_eval_scripts_end_to_end(tmp_path):
    g = np.load(GOLDEN[1])
    paths = {}
    for k in ('fx', 'moving', 'moved', 'seg_fx', 'seg_moving', 'seg_moved'):
        paths[k] = str(tmp_path / (k + '.nii.gz'))
        _nifti.save_nifti(g[k], paths[k], np.eye(4))
    mi = runpy.run_path(os.path.join(ROOT, 'scripts', 'eval_reg_with_mi.py'))
    out = str(tmp_path / 'nmi.csv')
    for _ in range(2):
        assert mi['main'](['--fx-im-path', paths['fx'], '--moving-im-path', paths['moving'][:-7], '--warped-im-path', paths['moved'],
                           '--sub-id', 'sub-07', '--out-file', out]) == 0
    rows = list(csv.reader(open(out)))
    assert rows[0] == ['Timestamp', 'Subject', 'NMI_before_registration', 'NMI_after_registration',
                       'NMI_between_moving_and_moved_images', 'Percentage_nmi_improvement_registration']
    assert len(rows) == 3 and rows[1][1] == 'sub-07'
    np.testing.assert_allclose([float(v) for v in rows[1][2:5]], [g['nmi_fx_moving'], g['nmi_fx_moved'], g['nmi_moving_moved']], rtol=1e-12)
    assert float(rows[1][5]) == np.round(100 * (g['nmi_fx_moved'] - g['nmi_fx_moving']) / g['nmi_fx_moving'], 2)

    sc = runpy.run_path(os.path.join(ROOT, 'scripts', 'eval_reg_on_sc_seg.py'))
    out = str(tmp_path / 'sc.csv')
    args = ['--fx-seg-path', paths['seg_fx'], '--moving-seg-path', paths['seg_moving'], '--warped-seg-path', paths['seg_moved'],
            '--sub-id', 'sub-07', '--out-file', out]
    assert sc['main'](args + ['--min-dice', '99', '--last-eval', '0']) == 1 and not os.path.exists(out)     # low Dice, not the last evaluation
    assert sc['main'](args + ['--min-dice', '99']) == 0
    rows = list(csv.reader(open(out)))
    assert rows[0][:4] == ['Timestamp', 'Subject', 'Dice_before_registration', 'Dice_after_registration'] and len(rows[0]) == 14
    got = [float(v) for v in rows[1][2:]]
    want = []
    for k in ('dice', 'jacc', 'sens', 'prec', 'spec', 'acc'):
        want += [g['sc_%s_moving' % k], g['sc_%s_moved' % k]]
    assert got == [float(v) for v in want]
