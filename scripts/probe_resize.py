import sys, torch
sys.path.insert(0, '.')
import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
B = 8
f = torch.randn(B, 3, 80, 80, 96, device='cuda').permute(0, 2, 3, 4, 1)
for _ in range(4):
    out = ops.rescale_dense_transform(f, 2)
torch.cuda.synchronize()
print('ok', out.shape)
