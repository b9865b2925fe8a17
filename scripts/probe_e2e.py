import sys, time, torch
sys.path.insert(0, '.')
import multimodal_registration_b200 as mrb
import bench
B = 32
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf, img = svf.pin_memory(), img.pin_memory()
model = mrb.voxelmorph.networks.VxmDense(bench.FULL, int_steps=7, svf_resolution=2, int_resolution=2)
for bs in (16, 8, 4, 2, 1):
    for _ in range(2): model.predict_deform([img, svf], copy=False, batch_size=bs)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): model.predict_deform([img, svf], copy=False, batch_size=bs)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
    print('batch_size %2d  %.2f ms  %.3e vox/s' % (bs, ms, B * bench.N_F / ms * 1e3))
