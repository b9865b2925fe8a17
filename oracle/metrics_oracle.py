"""CPU restatement (TEST INFRASTRUCTURE ONLY) of the reference's evaluation metrics:
  * normalized_mutual_information / detect_zero_padding -- eval_reg_with_mi.py:16-74
  * the overlap metrics of eval_reg_on_sc_seg.py:80-124
PINNED: tests/golden/metrics_*.npz hold the outputs of the reference's own lines executed here
(tests/golden/make_metrics_golden.py); tests/test_oracle_metrics.py checks this restatement against them."""
import numpy as np


def _entropy(pk):
    # scipy.stats.entropy with the default base (natural log)
    pk = np.asarray(pk, np.float64)
    pk = 1.0 * pk / np.sum(pk)
    vec = np.zeros_like(pk)
    nz = pk > 0
    vec[nz] = -pk[nz] * np.log(pk[nz])
    return np.sum(vec)


def joint_histogram(image0, image1, bins=100):
    hist, _ = np.histogramdd([np.reshape(image0, -1), np.reshape(image1, -1)], bins=bins)    # eval_reg_with_mi.py:66-69
    return hist


def normalized_mutual_information(image0, image1, bins=100):
    hist = joint_histogram(image0, image1, bins)
    h0 = _entropy(np.sum(hist, axis=0))                    # :71
    h1 = _entropy(np.sum(hist, axis=1))                    # :72
    h01 = _entropy(np.reshape(hist, -1))                   # :73
    return (h0 + h1) / h01


def detect_zero_padding(im):
    xy_plan = np.sum(im, axis=2)                           # :20-21
    yz_plan = np.sum(im, axis=0)
    x_plan = np.sum(xy_plan, axis=1)
    y_plan = np.sum(yz_plan, axis=1)
    z_plan = np.sum(yz_plan, axis=0)
    lo_hi = [(np.argwhere(p > 0)[0][0], np.argwhere(p > 0)[-1][0]) for p in (x_plan, y_plan, z_plan)]
    return lo_hi[0][0], lo_hi[1][0], lo_hi[2][0], lo_hi[0][1], lo_hi[1][1], lo_hi[2][1]


def overlap_metrics(fx, m):
    TP = np.sum(m[fx == 1])                                # eval_reg_on_sc_seg.py:80-86
    FP = np.sum(m[fx == 0])
    tn_tmp = m[fx == 0]
    TN = len(np.ravel(tn_tmp)) - np.sum(tn_tmp)
    fn_tmp = m[fx == 1]
    FN = len(np.ravel(fn_tmp)) - np.sum(fn_tmp)
    nb_vox, nb_sc = len(np.ravel(m)), np.sum(m)
    return dict(dice=(2 * TP) / (TP + TP + FP + FN), jaccard=TP / (TP + FP + FN), sensitivity=TP / (TP + FN),
                precision=TP / nb_sc, specificity=TN / (TN + FP), accuracy=(TP + TN) / nb_vox, TP=TP, FP=FP, TN=TN, FN=FN)
