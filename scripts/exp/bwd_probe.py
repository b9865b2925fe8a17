import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
from multimodal_registration_b200.ops import _ptr, _stream
import bench
for B in (1, 2, 8, 32):
    svf, _ = bench.synth_inputs(B, 'cpu', 0)
    v = ops.to_layout((svf / 64).cuda(), 'planar')        # |v| < 1: an early step
    g = ops.to_layout(torch.randn_like(svf).cuda(), 'planar')
    bound = torch.full((B,), float(v.abs().max()), device='cuda')
    gv = ops.empty(v.shape, 'planar', v.device)
    X, Y, Z = 80, 80, 96
    def run(bounded):
        if bounded: mrb._lib.call('dfm_ss_step_bwd_bounded', _ptr(g), _ptr(v), _ptr(gv), _ptr(bound), 1.0, B, X, Y, Z, 1.0, _stream())
        else: mrb._lib.call('dfm_ss_step_bwd', _ptr(g), _ptr(v), _ptr(gv), B, X, Y, Z, 1.0, _stream())
    for bounded in (False, True):
        for _ in range(3): run(bounded)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): run(bounded)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print('B=%2d %s: %.1f us  (%.2f of peak on 36 B/voxel)' % (B, 'gather ' if bounded else 'scatter', ms * 1e3, B * 36 * X * Y * Z / ms / 1e6 / 6504.1))
