"""Timing probe: image warp / SS step with an identity field vs the bench field (tuning aid)."""
import sys, torch
sys.path.insert(0, '.')
import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
import bench
B = 32
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()
flow = ops.rescale_dense_transform(ops.vecint(svf, 7), 2)
zero = torch.zeros_like(flow)
half = ops.vecint(svf, 6)
zero_h = torch.zeros_like(half)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
NF, NH = 160*160*192, 80*80*96
for name, f in [('bench flow', flow), ('identity', zero), ('const 3.3', zero + 3.3)]:
    ms = t(lambda: ops.warp(img, f))
    print('warp  %-12s %.3f ms  %.0f GB/s' % (name, ms, B*20*NF/ms/1e6))
for name, f in [('bench v6', half), ('identity', zero_h), ('const 0.3', zero_h + 0.3)]:
    ms = t(lambda: ops.vecint(f, 1))
    print('ss    %-12s %.3f ms  %.0f GB/s' % (name, ms, B*24*NH/ms/1e6))
ms = t(lambda: flow.clone())
print('clone full field %.3f ms %.0f GB/s' % (ms, 2*flow.numel()*4/ms/1e6))
