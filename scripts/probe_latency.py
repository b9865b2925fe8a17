"""Single-pair latency of the inference tail: eager launches vs CUDA-graph replay."""
import sys, time, torch
sys.path.insert(0, '.')
import multimodal_registration_b200 as mrb
import bench
svf, img = bench.synth_inputs(1, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()
model = mrb.voxelmorph.networks.VxmDense(bench.FULL, int_steps=7, svf_resolution=2, int_resolution=2)
def t(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
with torch.no_grad():
    e = t(lambda: model.deform([img, svf]))
    g = t(lambda: model.deform_graphed([img, svf]))
print('single pair (160x160x192): eager %.1f us  graph %.1f us  -> %.3e / %.3e vox/s' % (e, g, bench.N_F / e * 1e6, bench.N_F / g * 1e6))
