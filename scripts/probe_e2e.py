#!/usr/bin/env python
"""Tuning aid: PCIe copy bandwidths and the e2e call at several chunk sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import multimodal_registration_b200 as mrb

B = 32
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf_pin, img_pin = svf.pin_memory(), img.pin_memory()
dev = torch.device('cuda', 0)
d_img = torch.empty_like(img, device=dev)
h_out = torch.empty_like(img).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n

nb = img.numel() * 4
print('H2D  %.1f GB/s' % (nb / t(lambda: d_img.copy_(img_pin, non_blocking=True)) / 1e9))
print('D2H  %.1f GB/s' % (nb / t(lambda: h_out.copy_(d_img, non_blocking=True)) / 1e9))
def both():
    with torch.cuda.stream(s1): d_img.copy_(img_pin, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_img, non_blocking=True)
print('both %.1f GB/s each' % (nb / t(both) / 1e9))
vxm, _ = mrb.install_shims()
model = vxm.networks.VxmDense(bench.FULL, int_steps=bench.INT_STEPS, svf_resolution=2, int_resolution=2)
for bs in (1, 2, 4, 8, 16):
    dt = t(lambda: model.predict_deform([img_pin, svf_pin], copy=False, batch_size=bs), n=4)
    print('batch_size %2d: %.2f ms' % (bs, dt * 1e3))
