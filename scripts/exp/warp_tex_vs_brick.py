"""Tuning aid: stand-alone C=1 linear warp, texture-gather kernel (DFM_WARP_TEX=1) vs TMA-brick kernel, on fields of different
strength (bench field x 1, x 0.3, x 0; B=32, 160x160x192).  Run once per setting of DFM_WARP_TEX (read at library load)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
B = 32
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()
full = ops.rescale_dense_transform(ops.vecint(svf, 7), 2)


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


res = []
for scale in (1.0, 0.3, 0.0):
    f = full * scale
    f = ops.to_layout(f, 'planar')
    o = ops.warp(img, f)
    res.append('x%.1f: %.3f ms (checksum %.6f)' % (scale, timed(lambda: ops.warp(img, f)), float(o.double().sum())))
fcl = ops.to_layout(full, 'cl')
res.append('cl field: %.3f ms' % timed(lambda: ops.warp(img, fcl)))
res.append('fill: %.3f ms' % timed(lambda: ops.warp(img, full, fill_value=0.0)))
print('DFM_WARP_TEX=%s  ' % os.environ.get('DFM_WARP_TEX', '-') + ' | '.join(res))
