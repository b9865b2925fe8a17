"""Tuning aid: per-chunk timeline of VxmDense.predict_deform's three-stream pipeline (events on each stream)."""
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, multimodal_registration_b200 as mrb
from multimodal_registration_b200 import _host, ops
B = 32
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf_p, img_p = svf.pin_memory(), img.pin_memory()
model = mrb.voxelmorph.networks.VxmDense(bench.FULL, int_steps=7, svf_resolution=2, int_resolution=2)
for _ in range(3): model.predict_deform([img_p, svf_p], copy=False)
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = _host.device()
cur = torch.cuda.current_stream()
s_in, s_out = _host.side_streams()
y_host = _host.pinned_out((B,) + tuple(img.shape[1:]), torch.float32, 'out0')
for rep in range(2):
    torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True); ev0.record()
    evs = []
    t0 = time.perf_counter(); cpu = []
    for lo in range(0, B, bs):
        hi = lo + bs
        with torch.cuda.stream(s_in):
            a = torch.cuda.Event(enable_timing=True); a.record()
            src_d = img_p[lo:hi].to(dev, non_blocking=True); flow_d = svf_p[lo:hi].to(dev, non_blocking=True)
            b = torch.cuda.Event(enable_timing=True); b.record()
        cur.wait_stream(s_in)
        src_d.record_stream(cur); flow_d.record_stream(cur)
        c0 = torch.cuda.Event(enable_timing=True); c0.record()
        y, second = model.deform([src_d, flow_d], keep_pos_flow=False)
        y_c = ops.to_layout(y, 'cl')
        c1 = torch.cuda.Event(enable_timing=True); c1.record()
        s_out.wait_stream(cur)
        with torch.cuda.stream(s_out):
            d0 = torch.cuda.Event(enable_timing=True); d0.record()
            y_host[lo:hi].copy_(y_c, non_blocking=True)
            d1 = torch.cuda.Event(enable_timing=True); d1.record()
        y_c.record_stream(s_out)
        evs.append((a, b, c0, c1, d0, d1)); cpu.append(time.perf_counter() - t0)
    s_out.synchronize(); torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if rep == 1:
        print('batch_size', bs, 'wall %.2f ms' % (wall * 1e3), 'cpu loop done at %.2f ms' % (cpu[-1] * 1e3))
        for k, e in enumerate(evs):
            print('chunk %2d cpu %.2f | h2d %.2f-%.2f | compute %.2f-%.2f | d2h %.2f-%.2f' % ((k, cpu[k] * 1e3) + tuple(ev0.elapsed_time(x) for x in e)))

import cProfile, pstats
for bsz in (1, 2, 4):
    model.predict_deform([img_p, svf_p], copy=False, batch_size=bsz)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): model.predict_deform([img_p, svf_p], copy=False, batch_size=bsz)
    torch.cuda.synchronize(); print('predict_deform batch_size %d: %.2f ms' % (bsz, (time.perf_counter() - t0) / 5 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(3): model.predict_deform([img_p, svf_p], copy=False)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
