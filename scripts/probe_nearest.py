import sys, torch
sys.path.insert(0, '.')
import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
import bench
B = 8
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()
flow = ops.rescale_dense_transform(ops.vecint(svf, 7), 2)
seg = (img * 26).floor()
for _ in range(4):
    out = ops.warp(seg, flow, 'nearest', fill_value=0)
torch.cuda.synchronize()
print('ok')
