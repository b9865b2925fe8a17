"""Tuning aid: pred = warp(one_hot(labels, 26), flow) forward + d/d flow, generic channels-last kernels vs the label-map kernels
(B = 2, 160 x 160 x 192, the training-step shapes)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
B, C = 2, 26
svf, img = bench.synth_inputs(B, 'cpu', 0)
flow = ops.rescale_dense_transform(ops.vecint(svf.cuda() * 0.5, 5), 2)
g = torch.Generator(device='cpu').manual_seed(1)
labels = torch.randint(0, C, (B, 160, 160, 192), generator=g).cuda()
onehot = torch.nn.functional.one_hot(labels, C).float()
gpred = torch.rand(B, 160, 160, 192, C, generator=g).cuda()


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def fb(fn):
    f = flow.detach().requires_grad_(True)
    fn(f).backward(gpred)
    return f.grad


with torch.no_grad():
    a, b = ops.warp_onehot(labels, flow, C), ops.warp(onehot, flow)
    print('forward identical', bool(torch.equal(a, b)))
    print('forward: generic %.3f ms, label map %.3f ms' % (timed(lambda: ops.warp(onehot, flow)), timed(lambda: ops.warp_onehot(labels, flow, C))))
g1, g2 = fb(lambda f: ops.warp(onehot, f)), fb(lambda f: ops.warp_onehot(labels, f, C))
print('grad max |diff| %.3g (max |grad| %.3g)' % (float((g1 - g2).abs().max()), float(g1.abs().max())))
print('fwd+bwd: generic %.3f ms, label map %.3f ms' % (timed(lambda: fb(lambda f: ops.warp(onehot, f))), timed(lambda: fb(lambda f: ops.warp_onehot(labels, f, C)))))
