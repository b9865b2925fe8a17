"""One launch of the tld4 warp probe (for ncu)."""
import os, sys, ctypes, torch
here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(here)))
import bench, multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
B = int(os.environ.get('PROBE_B', 32)); nx = int(os.environ.get('PROBE_NX', 2)); mode = int(os.environ.get('PROBE_MODE', 0))
lib = ctypes.CDLL(os.path.join(here, 'libtexprobe.so'))
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()
full = ops.rescale_dense_transform(ops.vecint(svf, 7), 2)
fs = full.permute(0, 4, 1, 2, 3)
_, X, Y, Z, _ = full.shape
out = torch.empty((B, X, Y, Z, 1), device='cuda')
P = lambda t: ctypes.c_void_p(t.data_ptr())
lib.texprobe_setup(P(img), B, X, Y, Z)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    lib.texprobe_warp(P(fs), P(out), B, X, Y, Z, nx, mode, -1, st)
torch.cuda.synchronize()
