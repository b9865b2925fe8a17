"""Tuning aid: does the full-resolution field stay in the 126 MB L2 between the x2 up-sampler and the linear warp when the
batch is processed in chunks of `cs` pairs through ONE reused scratch buffer (59 MB per pair)?
  whole : U(all 32) W(all 32)                       -- the field makes a round trip through HBM (1.89 GB each way)
  slices: U(chunk) W(chunk) on slices of a full-size buffer   -- same launches, field still goes to HBM (control)
  reuse : U(chunk) W(chunk) through one chunk-sized scratch   -- dirty lines are overwritten in L2 before eviction
"""
import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops, _lib, _coords

B = int(os.environ.get('PROBE_B', 32))
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()
half = ops.vecint(svf, 7)                                  # planar [B,80,80,96,3]
_, Xh, Yh, Zh, _ = half.shape
X, Y, Z = 2 * Xh, 2 * Yh, 2 * Zh
dev = torch.cuda.current_device()
cx, cy, cz = (_coords.device_tables(n, 2 * n, dev)[0] for n in (Xh, Yh, Zh))
half_s = half.permute(0, 4, 1, 2, 3)                       # storage view [B,3,X,Y,Z]
assert half_s.is_contiguous()
out = torch.empty((B, X, Y, Z, 1), device='cuda')
full = torch.empty((B, 3, X, Y, Z), device='cuda')
P = ops._ptr
st = ops._stream


def U(src, dst, nb):
    _lib.call('dfm_resize_fwd', P(src), P(dst), P(cx), P(cy), P(cz), nb, 3, Xh, Yh, Zh, X, Y, Z, 2.0, 1.0,
              _lib.DFM_LINEAR, 0, st())


def W(im, fld, o, nb):
    _lib.call('dfm_warp_fwd', P(im), P(fld), P(o), nb, 1, X, Y, Z, X, Y, Z, _lib.DFM_LINEAR, 4, 0, 0.0, 0, 0, st())


def whole():
    U(half_s, full, B)
    W(img, full, out, B)


def chunked(cs, reuse, nscratch=1):
    scr = [torch.empty((cs, 3, X, Y, Z), device='cuda') for _ in range(nscratch)]

    def run():
        for k, b0 in enumerate(range(0, B, cs)):
            f = scr[k % nscratch] if reuse else full[b0:b0 + cs]
            U(half_s[b0:b0 + cs], f, cs)
            W(img[b0:b0 + cs], f, out[b0:b0 + cs], cs)
    return run


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


whole()
ref = out.clone()
print('whole            : %.3f ms' % timed(whole))
for cs in (1, 2, 4):
    for reuse in (False, True):
        fn = chunked(cs, reuse)
        out.zero_()
        fn()
        same = bool(torch.equal(out, ref))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        print('chunk %d %-6s   : %.3f ms eager, %.3f ms graph  (identical: %s)' % (
            cs, 'reuse' if reuse else 'slices', timed(fn), timed(g.replay), same))
