import sys, torch
sys.path.insert(0, '.')
import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
import bench
B, C = 2, 26
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf = svf.cuda()
lab = torch.randint(0, C, (B, 160, 160, 192), device='cuda')
oh = ops.to_layout(torch.nn.functional.one_hot(lab, C).float(), 'planar')
flow = ops.rescale_dense_transform(ops.vecint(svf, 5), 2)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
NF = 160*160*192
for name, f in [('bench flow', flow), ('identity', torch.zeros_like(flow)), ('0.3*flow', ops.to_layout(0.3 * ops.to_layout(flow, 'cl'), 'planar'))]:
    ms = t(lambda: ops.warp(oh, f))
    print('mc fwd %-10s %.3f ms %.0f GB/s' % (name, ms, B*(8*C+12)*NF/ms/1e6))
g = ops.to_layout(flow, 'cl')
print('flow grad stats: max |d/dz| %.3f  max|d/dx| %.3f' % ((g[:, :, :, 1:] - g[:, :, :, :-1]).abs().max().item(), (g[:, 1:] - g[:, :-1]).abs().max().item()))
