"""oracle/metrics_oracle.py against the goldens produced by executing the reference's own lines
(eval_reg_with_mi.py:16-74, eval_reg_on_sc_seg.py:80-124; tests/golden/make_metrics_golden.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import metrics_oracle as mo

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'metrics_*.npz')))


def crop(a, box):
    x0, y0, z0, x1, y1, z1 = [int(v) for v in box]
    return a[x0:x1 + 1, y0:y1 + 1, z0:z1 + 1]


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_metrics_oracle_matches_reference_outputs(path):
    g = np.load(path)
    assert len(GOLDEN) >= 3
    box = mo.detect_zero_padding(g['moving'])
    np.testing.assert_array_equal(np.array(box), g['box'])
    fx, moving, moved = (crop(g[k], box) for k in ('fx', 'moving', 'moved'))
    np.testing.assert_array_equal(mo.joint_histogram(fx, moved), g['hist_fx_moved'])
    for key, (a, b) in {'nmi_fx_moving': (fx, moving), 'nmi_fx_moved': (fx, moved), 'nmi_moving_moved': (moving, moved)}.items():
        np.testing.assert_allclose(mo.normalized_mutual_information(a, b), g[key], rtol=1e-13)
    np.testing.assert_allclose(mo.normalized_mutual_information(fx, moved, bins=7), g['nmi_bins7'], rtol=1e-13)
    for tag in ('moving', 'moved'):
        m = mo.overlap_metrics(g['seg_fx'], g['seg_' + tag])
        for ours, ref in (('dice', 'dice'), ('jaccard', 'jacc'), ('sensitivity', 'sens'), ('precision', 'prec'),
                          ('specificity', 'spec'), ('accuracy', 'acc'), ('TP', 'TP'), ('FP', 'FP'), ('TN', 'TN'), ('FN', 'FN')):
            assert m[ours] == g['sc_%s_%s' % (ref, tag)], (ours, tag)
