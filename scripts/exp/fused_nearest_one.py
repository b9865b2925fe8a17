import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
svf, img = bench.synth_inputs(32, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()
half = ops.vecint(svf, 7)
lab = (img * 25).round()
for _ in range(3):
    ops.rescale_warp(lab, half, 2, 0, 'nearest')
torch.cuda.synchronize()
