import sys, torch
sys.path.insert(0, '.')
import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
B, X, Y, Z, n = [int(a) for a in sys.argv[1:6]]
off = [float(a) for a in sys.argv[6:9]]
std = float(sys.argv[9])
torch.manual_seed(0)
c = torch.randn(B, X, Y, Z, 3, device='cuda') * std + torch.tensor(off, device='cuda')
flow = ops.vecint(ops.to_layout(c, 'planar'), n)
torch.cuda.synchronize()
ref = ops.vecint(c, n) if n <= 1 else None     # CL input + 1 step = direct kernel only
print('ok', None if ref is None else torch.equal(ops.to_layout(flow, 'cl'), ops.to_layout(ref, 'cl')))
