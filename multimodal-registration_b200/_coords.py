"""Host-side sampling-grid tables for ne.utils.resize (tf.linspace semantics, fp32).

The reference resamples onto ``tf.linspace(0., n_in - 1., n_out)`` per axis (corner aligned;
SURVEY.md Appendix A.2): delta = (n_in-1)/(n_out-1) in fp32, interior points
``0 + delta*k``, end points exact.  The tables are tiny (n_out floats per axis), computed here
once per (n_in, n_out) and cached on the device; the CUDA kernels take them as arguments so
the coordinate convention is a host decision.
"""
import functools

import numpy as np
import torch


def linspace_tf(n_in, n_out):
    start, stop = np.float32(0.), np.float32(n_in - 1)
    if n_out <= 0:
        return np.zeros((0,), np.float32)
    if n_out == 1:
        return np.array([start], np.float32)
    delta = np.float32(np.float32(stop - start) / np.float32(n_out - 1))
    k = np.arange(1, n_out - 1).astype(np.float32)
    inner = (start + (delta * k).astype(np.float32)).astype(np.float32)
    return np.concatenate([[start], inner, [stop]]).astype(np.float32)


def support_ranges(coords, n_in):
    """For each input index i: the [lo, hi) range of output indices whose two linear taps
    (clip(floor(c)), min(.+1, n_in-1)) include i.  coords is non-decreasing."""
    c = np.asarray(coords, np.float32)
    i0 = np.clip(np.floor(c), 0, n_in - 1).astype(np.int64)
    i1 = np.minimum(i0 + 1, n_in - 1)
    lo = np.zeros(n_in, np.int32)
    hi = np.zeros(n_in, np.int32)
    for i in range(n_in):
        hit = np.nonzero((i0 == i) | (i1 == i))[0]
        if hit.size:
            lo[i], hi[i] = hit[0], hit[-1] + 1
    return lo, hi


def adjoint_taps(coords, n_in):
    """(lo[n_in] int32, cnt[n_in] int32, w[n_in, kmax] float32): output j = lo[i] + k puts weight
    w[i, k] on input i (k < cnt[i]); weights follow the forward kernel's axis set-up exactly."""
    c = np.asarray(coords, np.float32)
    maxf = np.float32(n_in - 1)
    cl = np.clip(c, np.float32(0), maxf)
    i1 = np.minimum(np.floor(cl).astype(np.int64) + 1, n_in - 1)
    i0 = np.maximum(i1 - 1, 0)
    w0 = (i1.astype(np.float32) - cl).astype(np.float32)       # weight of the lower corner
    w1 = (np.float32(1) - w0).astype(np.float32)
    lo = np.zeros(n_in, np.int32)
    cnt = np.zeros(n_in, np.int32)
    taps = [[] for _ in range(n_in)]
    for j in range(len(c)):
        taps[i0[j]].append((j, w0[j]))
        if i1[j] != i0[j]:
            taps[i1[j]].append((j, w1[j]))
        else:
            taps[i1[j]][-1] = (j, np.float32(w0[j] + w1[j]))
    kmax = max(1, max(len(t) for t in taps))
    w = np.zeros((n_in, kmax), np.float32)
    for i, t in enumerate(taps):
        if not t:
            continue
        js = [a for a, _ in t]
        assert js == list(range(js[0], js[0] + len(js))), 'taps of a monotone grid are contiguous'
        lo[i], cnt[i] = js[0], len(js)
        w[i, :len(js)] = [b for _, b in t]
    return lo, cnt, w


@functools.lru_cache(maxsize=256)
def device_adjoint_taps(n_in, n_out, device_index):
    dev = torch.device('cuda', device_index)
    lo, cnt, w = adjoint_taps(linspace_tf(n_in, n_out), n_in)
    return (torch.from_numpy(lo).to(dev), torch.from_numpy(cnt).to(dev), torch.from_numpy(w).to(dev), int(w.shape[1]))


@functools.lru_cache(maxsize=256)
def device_tables(n_in, n_out, device_index):
    dev = torch.device('cuda', device_index)
    c = linspace_tf(n_in, n_out)
    lo, hi = support_ranges(c, n_in)
    return (torch.from_numpy(c).to(dev), torch.from_numpy(lo).to(dev), torch.from_numpy(hi).to(dev))
