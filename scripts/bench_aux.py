"""Secondary measurements (not the headline): Jacobian 256^3, compose, nearest warp, C=26 warp
fwd/bwd, VecInt bwd, rescale bwd.  Prints one JSON line per kernel with algorithmic GB/s and
the fraction of the measured HBM peak (same accounting as SURVEY.md section 8(d))."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
import bench

PEAK, _ = bench.measured_peak_gbs()


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def report(name, ms, nbytes, **kw):
    gbs = nbytes / ms / 1e6
    print(json.dumps(dict(kernel=name, ms=round(ms, 4), algorithmic_GB=round(nbytes / 1e9, 3), GBps=round(gbs, 1),
                          frac_of_peak=round(gbs / PEAK, 3), **kw)), flush=True)


def main():
    which = set(sys.argv[1:])
    torch.manual_seed(0)
    NF, NH = 160 * 160 * 192, 80 * 80 * 96
    if not which or 'jac' in which:
        B = 8
        f = torch.nn.functional.interpolate(torch.randn(B, 3, 16, 16, 16, device='cuda') * 4, size=(256, 256, 256),
                                            mode='trilinear').contiguous()           # planar storage
        field = f.permute(0, 2, 3, 4, 1)
        ms = timed(lambda: ops.jacobian_determinant(field, out_dtype=torch.float32))
        report('jacdet 256^3 f32->f32 (+stats)', ms, B * (12 * 256 ** 3 + 4 * 252 ** 3), B=B, vox_per_s=B * 252 ** 3 / ms * 1e3)
        ms = timed(lambda: ops.jacobian_determinant(field, out_dtype=torch.float64))
        report('jacdet 256^3 f32->f64 (+stats)', ms, B * (12 * 256 ** 3 + 8 * 252 ** 3), B=B)
        del f, field
    if not which or 'misc' in which:
        B = 32
        svf, img = bench.synth_inputs(B, 'cpu', 0)
        svf, img = svf.cuda(), img.cuda()
        a = ops.vecint(svf, 7)
        b = ops.vecint(-svf, 5)
        ms = timed(lambda: ops.compose([a, b]))
        report('compose half-res', ms, B * 36 * NH, B=B)
        flow = ops.rescale_dense_transform(a, 2)
        seg = (img * 26).floor()
        ms = timed(lambda: ops.warp(seg, flow, 'nearest', fill_value=0))
        report('warp nearest C=1 (fill 0)', ms, B * 20 * NF, B=B)
        ms = timed(lambda: ops.rescale_dense_transform(flow, 0.5))
        report('rescale x0.5', ms, B * (12 * NF + 12 * NH), B=B)
        ms = timed(lambda: ops.vecint(svf, 5))
        report('vecint 5 steps (cl in)', ms, B * 5 * 24 * NH, B=B)
        del a, b, flow, seg
    if not which or 'train' in which:
        B, C = 2, 26
        svf, img = bench.synth_inputs(B, 'cpu', 0)
        svf = svf.cuda()
        lab = torch.randint(0, C, (B, 160, 160, 192), device='cuda')
        onehot = torch.nn.functional.one_hot(lab, C).float()                  # channels-last like the reference
        flow = ops.rescale_dense_transform(ops.vecint(svf, 5), 2)
        for lay in ('cl', 'planar'):
            oh = ops.to_layout(onehot, lay)
            ms = timed(lambda: ops.warp(oh, flow), n=5)
            report('warp linear C=26 (%s image)' % lay, ms, B * (8 * C + 12) * NF, B=B)
            fl = flow.detach().clone().requires_grad_(True)
            pred = ops.warp(oh, fl)
            g = torch.rand_like(pred)
            def bwd():
                fl.grad = None
                pred.backward(g, retain_graph=True)
            ms = timed(bwd, n=5)
            report('warp bwd dfield C=26 (%s image)' % lay, ms, B * (8 * C + 24) * NF, B=B)
        s = svf.detach().clone().requires_grad_(True)
        out = ops.vecint(s, 5)
        g = torch.rand_like(out)
        def bwd2():
            s.grad = None
            out.backward(g, retain_graph=True)
        ms = timed(bwd2, n=5)
        report('vecint bwd 5 steps', ms, B * 5 * 36 * NH, B=B)
        h = ops.vecint(svf, 5).detach().requires_grad_(True)
        up = ops.rescale_dense_transform(h, 2)
        g = torch.rand_like(up)
        def bwd3():
            h.grad = None
            up.backward(g, retain_graph=True)
        ms = timed(bwd3, n=5)
        report('rescale x2 bwd', ms, B * (12 * NF + 12 * NH), B=B)

    if not which or 'loss' in which:
        B, C = 2, 26
        lab = torch.randint(0, C, (B, 160, 160, 192), device='cuda')
        t = torch.nn.functional.one_hot(lab, C).float()
        p = torch.rand_like(t).requires_grad_(True)
        ms = timed(lambda: ops.dice_loss(t, p.detach()), n=5)
        report('dice sums C=26 (cl)', ms, B * 2 * 4 * C * NF, B=B)
        loss = ops.dice_loss(t, p)
        def bwd4():
            p.grad = None
            loss.backward(retain_graph=True)
        ms = timed(bwd4, n=5)
        report('dice bwd C=26 (cl)', ms, B * 2 * 4 * C * NF, B=B)
        f = torch.randn(B, 160, 160, 192, 3, device='cuda').requires_grad_(True)
        ms = timed(lambda: ops.grad_l2_loss(f.detach(), 0.5), n=5)
        report('grad l2 sums (cl field)', ms, B * 12 * NF, B=B)
        gl = ops.grad_l2_loss(f, 0.5).sum()
        def bwd5():
            f.grad = None
            gl.backward(retain_graph=True)
        ms = timed(bwd5, n=5)
        report('grad l2 bwd (cl field)', ms, B * 24 * NF, B=B)

    if not which or 'fused_dice' in which:
        B, C = 2, 26
        svf, _ = bench.synth_inputs(B, 'cpu', 0)
        flow = ops.rescale_dense_transform(ops.vecint(svf.cuda(), 5), 2)
        moving = torch.nn.functional.one_hot(torch.randint(0, C, (B, 160, 160, 192), device='cuda'), C).float()
        fixed = torch.nn.functional.one_hot(torch.randint(0, C, (B, 160, 160, 192), device='cuda'), C).float()
        nbytes = B * ((8 * C + 12) * NF + (8 * C + 24) * NF)          # warp fwd + d/dfield, the unfused accounting
        def two_ops():
            fl = flow.detach().requires_grad_(True)
            ops.dice_loss(fixed, ops.warp(moving, fl)).backward()
            return fl.grad
        def one_op():
            fl = flow.detach().requires_grad_(True)
            ops.warp_dice_loss(moving, fl, fixed).backward()
            return fl.grad
        ms = timed(two_ops, n=5)
        report('warp + Dice fwd+bwd C=26, stand-alone ops', ms, nbytes, B=B)
        ms = timed(one_op, n=5)
        report('warp + Dice fwd+bwd C=26, fused', ms, nbytes, B=B)


if __name__ == '__main__':
    main()
