"""Driver of scripts/exp/tex_probe.cu: texture-gather (tld4) linear warp against the library's TMA-brick warp on the bench field."""
import os, sys, ctypes, torch
here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(here)))
import bench, multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops

B = int(os.environ.get('PROBE_B', 32))
lib = ctypes.CDLL(os.path.join(here, 'libtexprobe.so'))
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()
full = ops.rescale_dense_transform(ops.vecint(svf, 7), 2)          # planar [B,160,160,192,3]
fs = full.permute(0, 4, 1, 2, 3)
assert fs.is_contiguous()
_, X, Y, Z, _ = full.shape
ref = ops.warp(img, full)
out = torch.empty_like(ref)
P = lambda t: ctypes.c_void_p(t.data_ptr())
rc = lib.texprobe_setup(P(img), B, X, Y, Z)
print('setup rc', rc)
if rc:
    sys.exit(0)


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


print('library warp: %.3f ms' % timed(lambda: ops.warp(img, full)))
for nx in (2, 4):
    for mode in (0, 1, 2, 3, 4, 7):
        for carve in (-1, 0):
            out.zero_()
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            fn = lambda: lib.texprobe_warp(P(fs), P(out), B, X, Y, Z, nx, mode, carve, st)
            rc = fn()
            torch.cuda.synchronize()
            print('tld4 warp NX=%d mode=%d carve=%d: rc %d, %.3f ms, identical %s' % (nx, mode, carve, rc, timed(fn), bool(torch.equal(out, ref))))
