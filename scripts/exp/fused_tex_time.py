"""Time of the fused texture-gather kernel (variant / carve-out from the environment: DFM_TEX_VARIANT, DFM_TEX_CARVEOUT)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
B = int(os.environ.get('PROBE_B', 32))
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()
half = ops.vecint(svf, 7)
ref = ops.warp(img, ops.rescale_dense_transform(half, 2))
out = ops.rescale_warp(img, half, 2)
ok = bool(torch.equal(out, ref))
for _ in range(3):
    ops.rescale_warp(img, half, 2)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    ops.rescale_warp(img, half, 2)
b.record()
torch.cuda.synchronize()
print('hyb %s carve %s: %.3f ms identical %s' % (os.environ.get('DFM_TEX_HYB', '-'), os.environ.get('DFM_TEX_CARVEOUT', '-'), a.elapsed_time(b) / 20, ok))
