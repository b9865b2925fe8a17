"""Tuning aid: the e2e leg of bench.py alone (VxmDense.predict_deform on pinned host arrays, B = 32)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, multimodal_registration_b200 as mrb
svf, img = bench.synth_inputs(32, 'cpu', 0)
svf_p, img_p = svf.pin_memory(), img.pin_memory()
model = mrb.voxelmorph.networks.VxmDense(bench.FULL, int_steps=7, svf_resolution=2, int_resolution=2)
for _ in range(3):
    model.predict_deform([img_p, svf_p], copy=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    model.predict_deform([img_p, svf_p], copy=False)
torch.cuda.synchronize()
print('e2e %.2f ms per step (NO_MEAS=%s)' % ((time.perf_counter() - t0) / 5 * 1e3, os.environ.get('DFM_MARCH_NO_MEAS')))
# device time of one chunk of 2 pairs
s2, i2 = svf[:2].cuda(), img[:2].cuda()
for _ in range(3):
    model.deform([i2, s2], keep_pos_flow=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    model.deform([i2, s2], keep_pos_flow=False)
b.record()
torch.cuda.synchronize()
print('chunk of 2 pairs on the device: %.3f ms' % (a.elapsed_time(b) / 10))
