import torch
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
n = 1 << 29   # 2 GiB of fp32
x = torch.empty(n, device='cuda'); y = torch.empty(n, device='cuda')
ms = t(lambda: x.fill_(1.0)); print('fill (write only)  %.3f ms  %.0f GB/s' % (ms, n*4/ms/1e6))
ms = t(lambda: x.sum());      print('sum  (read only)   %.3f ms  %.0f GB/s' % (ms, n*4/ms/1e6))
ms = t(lambda: y.copy_(x));   print('copy (read+write)  %.3f ms  %.0f GB/s' % (ms, 2*n*4/ms/1e6))
z = torch.empty(n // 8, device='cuda')
ms = t(lambda: x.view(8, -1).copy_(z.expand(8, -1)));   print('1 read : 8 write     %.3f ms  %.0f GB/s' % (ms, (n + n//8)*4/ms/1e6))
