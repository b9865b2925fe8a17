"""Voxel-level evaluation metrics of the reference's eval scripts, on the device.

``normalized_mutual_information`` / ``detect_zero_padding`` keep the names and argument meaning of
eval_reg_with_mi.py:16-74; ``overlap_metrics`` returns the quantities eval_reg_on_sc_seg.py:80-124 derives
from two segmentations.  The voxel passes (min / max, joint histogram, plane sums, masked sums) are CUDA
kernels behind the C ABI (dfm.h: dfm_minmax, dfm_joint_hist, dfm_axis_sums, dfm_overlap_sums); the handful of
scalars that follow (entropies of a 100 x 100 table, ratios) are formed on the host in float64 like the
reference does.  Inputs: numpy arrays or tensors, float32 or float64 (``get_fdata()`` gives float64)."""
import numpy as np
import torch

from . import _host, _lib
from .ops import _ptr, _stream


def _as_device(a):
    t = _host.to_device(a, tag='metric')
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)                       # get_fdata() semantics: everything becomes float64
    return t.contiguous()


def _pair(a, b):
    a, b = _as_device(a), _as_device(b)
    if a.dtype != b.dtype:
        a, b = a.to(torch.float64), b.to(torch.float64)
    if a.numel() != b.numel():
        raise ValueError('the two images must have the same number of elements (got %d and %d)' % (a.numel(), b.numel()))
    return a, b


def _work(dev):
    return torch.empty(max(_lib.load().dfm_metrics_workspace_bytes() // 8, 1), device=dev, dtype=torch.float64)


def joint_histogram(image0, image1, bins=100):
    """``np.histogramdd([image0.ravel(), image1.ravel()], bins=bins)[0]`` as a uint64-exact int64 tensor
    [bins, bins] on the device (row = bin of image0).  float32 inputs are binned as their float64 values (the
    reference scripts only ever see ``get_fdata()`` float64 arrays): edges and comparisons are float64."""
    a, b = _pair(image0, image1)
    is64 = int(a.dtype == torch.float64)
    mm = torch.empty((2, 2), device=a.device, dtype=torch.float64)
    work = _work(a.device)
    _lib.call('dfm_minmax', _ptr(a), a.numel(), is64, _ptr(mm[0]), _ptr(work), _stream())
    _lib.call('dfm_minmax', _ptr(b), b.numel(), is64, _ptr(mm[1]), _ptr(work), _stream())
    hist = torch.empty((bins, bins), device=a.device, dtype=torch.int64)
    _lib.call('dfm_joint_hist', _ptr(a), _ptr(b), a.numel(), is64, _ptr(mm[0]), _ptr(mm[1]), int(bins), _ptr(hist), _stream())
    return hist


def _entropy(pk):
    # scipy.stats.entropy(pk): pk / sum(pk), then sum(entr(pk)) with entr(x) = -x log x, entr(0) = 0, natural log
    pk = np.asarray(pk, np.float64)
    pk = 1.0 * pk / np.sum(pk, axis=0, keepdims=True)
    with np.errstate(divide='ignore', invalid='ignore'):
        vec = np.where(pk > 0, -pk * np.log(pk), 0.0)
    return np.sum(vec, axis=0)


def normalized_mutual_information(image0, image1, bins=100):
    """eval_reg_with_mi.py:38-74 (scikit-image's NMI, Studholme et al.): (H0 + H1) / H01 of the joint histogram."""
    hist = joint_histogram(image0, image1, bins).cpu().numpy().astype(np.float64)
    h0 = _entropy(np.sum(hist, axis=0))
    h1 = _entropy(np.sum(hist, axis=1))
    h01 = _entropy(np.reshape(hist, -1))
    return float((h0 + h1) / h01)


def detect_zero_padding(im):
    """eval_reg_with_mi.py:16-36: (x_min, y_min, z_min, x_max, y_max, z_max) of the planes whose sum is > 0."""
    t = _as_device(im)
    if t.dim() != 3:
        raise ValueError('detect_zero_padding: a 3-D volume is expected, got shape %s' % (tuple(t.shape),))
    X, Y, Z = t.shape
    sums = torch.empty(X + Y + Z, device=t.device, dtype=torch.float64)
    _lib.call('dfm_axis_sums', _ptr(t), X, Y, Z, int(t.dtype == torch.float64), _ptr(sums[:X]), _ptr(sums[X:X + Y]),
              _ptr(sums[X + Y:]), _stream())
    s = sums.cpu().numpy()
    out = []
    for plan in (s[:X], s[X:X + Y], s[X + Y:]):
        nz = np.argwhere(plan > 0)
        out.append((int(nz[0][0]), int(nz[-1][0])))         # IndexError on an all-zero volume, like the reference
    return out[0][0], out[1][0], out[2][0], out[0][1], out[1][1], out[2][1]


def overlap_counts(fx_seg, seg):
    """TP, FP, TN, FN and the voxel counts of eval_reg_on_sc_seg.py:80-98 for one (fixed, other) pair of
    segmentations, float64 like the reference's numpy sums."""
    f, m = _pair(fx_seg, seg)
    out = torch.empty(5, device=f.device, dtype=torch.float64)
    _lib.call('dfm_overlap_sums', _ptr(f), _ptr(m), f.numel(), int(f.dtype == torch.float64), _ptr(out), _ptr(_work(f.device)), _stream())
    s1, s0, n1, n0, sm = [np.float64(v) for v in out.cpu().numpy()]
    return dict(TP=s1, FP=s0, TN=np.float64(n0 - s0), FN=np.float64(n1 - s1), nb_vox=int(f.numel()), nb_sc_vox=sm)


def overlap_metrics(fx_seg, seg):
    """Dice, Jaccard, sensitivity, precision, specificity, accuracy as eval_reg_on_sc_seg.py:100-124 forms them."""
    c = overlap_counts(fx_seg, seg)
    TP, FP, TN, FN = c['TP'], c['FP'], c['TN'], c['FN']
    with np.errstate(divide='ignore', invalid='ignore'):
        return dict(dice=(2 * TP) / (TP + TP + FP + FN), jaccard=TP / (TP + FP + FN), sensitivity=TP / (TP + FN),
                    precision=TP / c['nb_sc_vox'], specificity=TN / (TN + FP), accuracy=(TP + TN) / c['nb_vox'], **c)
