"""Tuning aid: does the x2 up-sampler (DRAM-write bound) overlap with the linear warp (LSU bound) when the batch is split in
halves on two streams?  Sequential: U(all) W(all).  Overlapped: U(A) | W(A) || U(B) | W(B)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
B = 32
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()
half = ops.vecint(svf, 7)
def seq():
    return ops.warp(img, ops.rescale_dense_transform(half, 2))
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print('sequential U+W: %.3f ms' % timed(seq))
for nch in (2, 4):
    for prio in (0, -1):
        s_hi = torch.cuda.Stream(priority=prio)
        cur = torch.cuda.current_stream()
        cs = B // nch
        def ovl():
            outs = []
            s_hi.wait_stream(cur)
            ups = [None] * nch
            evs = [torch.cuda.Event() for _ in range(nch)]
            with torch.cuda.stream(s_hi):
                for k in range(nch):
                    ups[k] = ops.rescale_dense_transform(half[k * cs:(k + 1) * cs], 2)
                    evs[k].record(s_hi)
            for k in range(nch):
                cur.wait_event(evs[k])
                outs.append(ops.warp(img[k * cs:(k + 1) * cs], ups[k]))
                ups[k].record_stream(cur)
            return outs
        print('chunks %d, up-sampler stream priority %d: %.3f ms' % (nch, prio, timed(ovl)))
