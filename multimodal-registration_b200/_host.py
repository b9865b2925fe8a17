"""Host <-> device plumbing for the Keras-style entry points (numpy in, numpy out).

Inputs are staged through cached PINNED buffers and copied with cudaMemcpyAsync on the current
stream; results come back through pinned buffers too.  This is the path `bench.py` times as
`e2e` (host buffers, copies inside the timed region).
"""
import numpy as np
import torch

import collections
import os

_PINNED = collections.OrderedDict()      # key -> pinned staging buffer, least recently used first
_BUSY = {}      # staging buffer key -> event of the last async copy that READ from it
# Pinned host memory is a scarce resource: the cache is bounded by BYTES and evicts least-recently-used
# buffers (BIDS runs see a different shape per subject).  Lifetime of `copy=False` views: a view handed out
# by to_host(copy=False) / predict(copy=False) stays valid while its buffer is cached; an evicted buffer is
# only dropped from the cache -- the numpy view keeps the pinned allocation alive -- but a later call with
# the same (shape, dtype, tag) REUSES the buffer and overwrites the view.
PINNED_CACHE_BYTES = int(os.environ.get('DFM_PINNED_CACHE_MB', '4096')) << 20


def _nbytes(buf):
    return buf.numel() * buf.element_size()


def _pinned(shape, dtype, tag):
    key = (tuple(shape), dtype, tag)
    buf = _PINNED.get(key)
    if buf is None:
        buf = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
        _PINNED[key] = buf
        total = sum(_nbytes(b) for b in _PINNED.values())
        while total > PINNED_CACHE_BYTES and len(_PINNED) > 1:
            old_key, old = next(iter(_PINNED.items()))
            if old_key == key:
                break
            ev = _BUSY.pop(old_key, None)
            if ev is not None:
                ev.synchronize()
            del _PINNED[old_key]
            total -= _nbytes(old)
    else:
        _PINNED.move_to_end(key)
    ev = _BUSY.get(key)
    if ev is not None:
        ev.synchronize()          # an earlier H2D copy may still be reading this buffer
    return buf


_POOL = None
COPY_THREADS = int(os.environ.get('DFM_HOST_COPY_THREADS', str(min(8, os.cpu_count() or 1))))


def parallel_copy_(dst, src):
    """dst.copy_(src) for large contiguous CPU tensors of one dtype, split over a few threads: a single-threaded
    memcpy (~8 GB/s) is what limits the pageable-numpy entry (nibabel arrays in, fresh arrays out), not PCIe."""
    global _POOL
    n = dst.numel()
    if (COPY_THREADS <= 1 or n * dst.element_size() < (32 << 20) or dst.dtype != src.dtype
            or not dst.is_contiguous() or not src.is_contiguous()):
        dst.copy_(src)
        return dst
    if _POOL is None:
        import concurrent.futures
        _POOL = concurrent.futures.ThreadPoolExecutor(COPY_THREADS)
    d, s = dst.view(-1), src.view(-1)
    step = -(-n // COPY_THREADS)
    futs = [_POOL.submit(d[i:i + step].copy_, s[i:i + step]) for i in range(0, n, step)]     # copy_ releases the GIL
    for f in futs:
        f.result()
    return dst


def host_copy(stage):
    """Fresh (pageable) numpy copy of a pinned staging tensor."""
    out = torch.empty(stage.shape, dtype=stage.dtype)
    return parallel_copy_(out, stage).numpy()


def _mark_busy(shape, dtype, tag):
    ev = torch.cuda.Event()
    ev.record()
    _BUSY[(tuple(shape), dtype, tag)] = ev


def device():
    if not torch.cuda.is_available():
        from ._lib import DfmError
        raise DfmError('no CUDA device: the deformation engine has no CPU path')
    return torch.device('cuda', torch.cuda.current_device())


def to_device(x, dtype=None, tag='in'):
    """numpy array / CPU tensor / CUDA tensor -> CUDA tensor (optionally cast to dtype)."""
    dev = device()
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            return x if dtype is None or x.dtype == dtype else x.to(dtype)
        t = x
    else:
        a = np.asarray(x)
        if not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        if a.dtype == np.float64 and dtype == torch.float32:
            a = a.astype(np.float32)          # Keras casts inputs to floatx on the host too
        t = torch.from_numpy(a)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_pinned():
        stage = _pinned(t.shape, t.dtype, tag)
        parallel_copy_(stage, t)
        out = stage.to(dev, non_blocking=True)
        _mark_busy(t.shape, t.dtype, tag)   # the staging buffer must not be overwritten before this copy ran
        return out
    return t.to(dev, non_blocking=True)


def to_host(t, tag='out', copy=True):
    """CUDA tensor (any physical layout) -> C-contiguous numpy array of its logical shape.
    copy=False returns a view of the cached pinned staging buffer (valid until the next call
    that produces an output of the same shape and tag) and saves one host memcpy."""
    if not t.is_contiguous():
        from . import ops
        t = ops.to_layout(t, 'cl')
    stage = _pinned(t.shape, t.dtype, tag)
    stage.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host_copy(stage) if copy else stage.numpy()


_STREAMS = {}


def side_streams():
    """Two cached non-default streams per device (H2D copies, D2H copies)."""
    idx = torch.cuda.current_device()
    if idx not in _STREAMS:
        _STREAMS[idx] = (torch.cuda.Stream(), torch.cuda.Stream())
    return _STREAMS[idx]


def pinned_view(x, dtype, tag):
    """Host array / tensor -> pinned CPU tensor of ``dtype`` (no copy if it already is one)."""
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    if t.dtype != dtype:
        t = t.to(dtype)
    if t.is_pinned() and t.is_contiguous():
        return t
    stage = _pinned(t.shape, t.dtype, tag)
    parallel_copy_(stage, t)
    return stage


def pinned_out(shape, dtype, tag):
    return _pinned(shape, dtype, tag)
