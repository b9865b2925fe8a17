python -m pytest tests/test_metrics_gpu.py tests/test_gpu_parity.py -x -q -k "metrics or histogram or overlap or eval_scripts or channelwise" 2>&1 | tail -15
python - <<'PY'
import sys,time,torch
sys.path.insert(0,'.')
import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops, metrics
# channel-wise warp at the reference size: [160,160,192,26] with a smooth per-channel field
torch.manual_seed(0)
X,Y,Z,C=160,160,192,26
vol=torch.rand(1,X,Y,Z,C,device='cuda')
coarse=torch.randn(1,3*2,10,10,12,device='cuda')*3   # 2 samples along the label axis (ceil(26/16))
f=torch.nn.functional.interpolate(coarse,size=(X,Y,Z),mode='trilinear',align_corners=True)  # [1,6,X,Y,Z]
f=f.reshape(1,2,3,X,Y,Z)
w=torch.linspace(0,1,C,device='cuda').view(1,C,1,1,1,1)
field=(f[:,0:1]*(1-w)+f[:,1:2]*w).permute(0,3,4,5,1,2).contiguous()   # [1,X,Y,Z,C,3]
del f
def timed(fn,n=5):
    fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n
N=X*Y*Z
ms=timed(lambda: ops.warp_channelwise(vol,field))
print('channelwise fwd %.3f ms  %.1f GB/s (20C B/voxel)'%(ms, N*C*20/ms/1e6))
ms=timed(lambda: ops.warp_channelwise(vol,field,argmax=True))
print('channelwise+argmax %.3f ms  %.1f GB/s (16C+1 B/voxel)'%(ms, N*(C*16+1)/ms/1e6))
a=torch.rand(160,160,192,device='cuda'); b=0.5*a+0.5*torch.rand_like(a)
ms=timed(lambda: metrics.joint_histogram(a,b))
print('joint hist 160x160x192 f32: %.3f ms (%.1f GB/s incl. 2 min/max passes)'%(ms, N*16/ms/1e6))
PY
