import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops, _coords
from multimodal_registration_b200.ops import _ptr, _stream
Xi, Yi, Zi, Xo, Yo, Zo, C = 80, 80, 96, 160, 160, 192, 3
t = [_coords.device_adjoint_taps(a, b, 0) for a, b in ((Xi, Xo), (Yi, Yo), (Zi, Zo))]
for B in (2, 16):
    gout = ops.empty((B, Xo, Yo, Zo, C), 'planar', torch.device('cuda')); gout.normal_()
    gin = ops.empty((B, Xi, Yi, Zi, C), 'planar', gout.device)
    work = torch.empty(mrb._lib.load().dfm_resize_bwd_workspace_bytes(B, C, Xi, Yi, Zo) // 4, device='cuda')
    for ws in (None, work):
        def run():
            mrb._lib.call('dfm_resize_bwd_ws', _ptr(gout), _ptr(gin), _ptr(ws),
                          _ptr(t[0][0]), _ptr(t[0][1]), _ptr(t[0][2]), t[0][3], _ptr(t[1][0]), _ptr(t[1][1]), _ptr(t[1][2]), t[1][3],
                          _ptr(t[2][0]), _ptr(t[2][1]), _ptr(t[2][2]), t[2][3], B, C, Xi, Yi, Zi, Xo, Yo, Zo, 2.0, 1.0, _stream())
        for _ in range(3): run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): run()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print('B=%2d %s: %.1f us (%.2f of peak on 12(N_in+N_out) B)' % (B, 'separable' if ws is not None else 'one-pass ', ms * 1e3,
              B * 12 * (Xi * Yi * Zi + Xo * Yo * Zo) / ms / 1e6 / 6504.1))
