"""Tuning aid: fused rescale + nearest warp against the two stand-alone kernels (B=32, bench shapes)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
B = 32
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()
half = ops.vecint(svf, 7)
lab = (img * 25).round()


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


ref = ops.warp(lab, ops.rescale_dense_transform(half, 2), 'nearest', 0)
out = ops.rescale_warp(lab, half, 2, 0, 'nearest')
print('identical', bool(torch.equal(ref, out)))
print('nearest: two kernels %.3f ms, fused %.3f ms' % (timed(lambda: ops.warp(lab, ops.rescale_dense_transform(half, 2), 'nearest', 0)),
                                                      timed(lambda: ops.rescale_warp(lab, half, 2, 0, 'nearest'))))
