"""Tuning aid: fused rescale+warp through the texture-gather kernel (dfm_rescale_warp_fwd) against the two stand-alone
kernels, both builds: identity of the results and time at the bench shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops, _lib

B = int(os.environ.get('PROBE_B', 32))
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()


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


for exact in (False, True):
    _lib.use(exact)
    half = ops.vecint(svf, 7)
    ref = ops.warp(img, ops.rescale_dense_transform(half, 2))
    out = ops.rescale_warp(img, half, 2)
    torch.cuda.synchronize()
    print('exact=%s identical=%s max|diff|=%.3g' % (exact, bool(torch.equal(out, ref)), (out - ref).abs().max().item()))
    reff = ops.warp(img, ops.rescale_dense_transform(half, 2), fill_value=-1.0)
    outf = ops.rescale_warp(img, half, 2, fill_value=-1.0)
    print('   fill: identical=%s n_fill=%d' % (bool(torch.equal(outf, reff)), int((outf == -1.0).sum())))
    print('   two kernels %.3f ms, fused %.3f ms' % (timed(lambda: ops.warp(img, ops.rescale_dense_transform(half, 2))),
                                                    timed(lambda: ops.rescale_warp(img, half, 2))))
