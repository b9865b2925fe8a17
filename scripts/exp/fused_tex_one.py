"""A few launches of the fused texture-gather kernel (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
B = int(os.environ.get('PROBE_B', 32))
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()
half = ops.vecint(svf, 7)
for _ in range(3):
    out = ops.rescale_warp(img, half, 2)
torch.cuda.synchronize()
