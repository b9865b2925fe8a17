import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
import bench
# correctness: compare cfg kernel against DFM_WARP_DIRECT? -> compare with oracle on a small rough field
from oracle import interp_oracle as io
rng = np.random.default_rng(0)
for shape in ((24, 20, 48), (17, 9, 36), (8, 8, 16)):
    img = rng.random((2,) + shape + (1,)).astype(np.float32)
    f = rng.standard_normal((2,) + shape + (3,)).astype(np.float32)
    for ax in (1, 2, 3):
        f = (f + np.roll(f, 1, ax) + np.roll(f, -1, ax)) / 3
    f = (f / f.std() * 2.5).astype(np.float32)
    for fill in (None, 0.0):
        want = io.spatial_transformer(img, f, 'linear', fill)
        for lay in ('planar', 'cl'):
            got = ops.to_layout(ops.warp(torch.from_numpy(img).cuda(), ops.to_layout(torch.from_numpy(f).cuda(), lay), 'linear', fill), 'cl').cpu().numpy()
            print(shape, fill, lay, 'max err', np.abs(got - want).max())
svf, img = bench.synth_inputs(8, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()
flow = ops.rescale_dense_transform(ops.vecint(svf, 7), 2)
for _ in range(5): out = ops.warp(img, flow)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): out = ops.warp(img, flow)
e1.record(); torch.cuda.synchronize()
print('CFG', os.environ.get('DFM_WARP_CFG'), 'warp ms at B=8: %.4f  (x4 = %.3f)' % (e0.elapsed_time(e1) / 20, e0.elapsed_time(e1) / 5))
