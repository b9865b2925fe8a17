"""GPU parity at the BASELINE shapes and in the regime bench.py runs.

The small-volume tests of test_gpu_parity.py never reach the code paths the benchmark takes: the
plane-marching SS kernel (Z >= 33), its halo-2 / halo-4 variant selection and global-gather fallback, the
image brick's fit / no-fit tiles on a std-3 field, 32-bit voxel offsets near 2^23, the 256^3 Jacobian.
Here the CUDA path (through the C ABI) is compared with the CPU oracle on `bench.synth_inputs` itself at
80x80x96 -> 160x160x192, in both builds: exact build bit-for-bit, default build within the north_star bar
(1e-5 relative / 1e-4 voxel absolute) with the observed maximum error printed (run with -s; the numbers
are also recorded in DESIGN.md).  Oracle results are cached per module so both builds share them.
"""
import functools
import os
import sys

import numpy as np
import pytest
import torch

import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
from oracle import interp_oracle as io
from oracle import jacobian_oracle as jo
from oracle import torch_oracle as to

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synth_inputs: the benchmark's own field / image generator)

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-4          # north_star tolerance for fp32 linear paths


@pytest.fixture(autouse=True, params=['fast', 'exact'])
def arithmetic_mode(request):
    mrb._lib.use(request.param == 'exact')
    yield request.param
    mrb._lib.use(False)


def dev(a, layout='cl'):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return ops.to_layout(t, layout) if t.dim() >= 3 else t


def host(t):
    return ops.to_layout(t, 'cl').cpu().numpy() if t.dim() >= 3 else t.cpu().numpy()


def report(name, got, want):
    """Parity bar of the north_star + the observed maximum errors (stdout, -s)."""
    g, w = got.astype(np.float64), want.astype(np.float64)
    err = np.abs(g - w)
    rel = err / np.maximum(np.abs(w), 1e-30)
    # the bar is |got - want| <= ATOL + RTOL |want|: report the worst excess ratio too
    ratio = (err / (ATOL + RTOL * np.abs(w))).max()
    print('\n[parity %s | %s build] max abs err %.3e, max rel err (|want| > 1e-3) %.3e, worst err/(atol+rtol|want|) %.3e'
          % (name, 'exact' if mrb._lib.exact_order() else 'default', err.max(),
             rel[np.abs(w) > 1e-3].max() if (np.abs(w) > 1e-3).any() else 0.0, ratio))
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)
    if mrb._lib.exact_order():
        np.testing.assert_array_equal(got, want)


# --------------------------------------------------------------------------------------
# oracle side, computed once per session (both builds compare against the same arrays)
# --------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def headline_oracle():
    torch.set_num_threads(os.cpu_count() or 1)
    svf, img = bench.synth_inputs(1, 'cpu', 0)
    svf, img = svf.numpy(), img.numpy()
    flow_h = io.vec_int(svf, 7)                               # [1, 80, 80, 96, 3]
    flow_f = io.rescale_dense_transform(flow_h, 2)            # [1, 160, 160, 192, 3]
    moved = io.spatial_transformer(img, flow_f, 'linear')
    labels = (img * 26).astype(np.uint8)                      # label map for the nearest path
    moved_nn = io.spatial_transformer(labels, flow_f, 'nearest', 0)
    return svf, img, flow_h, flow_f, moved, labels, moved_nn


def test_headline_pipeline_at_baseline_shape():
    """bench.synth_inputs (std-3 SVF) through VecInt(7) @80x80x96 -> x2 -> linear + nearest warp @160x160x192."""
    svf, img, flow_h, flow_f, moved, labels, moved_nn = headline_oracle()
    d_svf, d_img = dev(svf), dev(img)
    g_flow_h = ops.vecint(d_svf, 7)
    report('VecInt(7) 80x80x96', host(g_flow_h), flow_h)
    # planar svf takes the planar first-step instantiation
    report('VecInt(7) 80x80x96 planar svf', host(ops.vecint(ops.to_layout(d_svf, 'planar'), 7)), flow_h)
    g_flow_f = ops.rescale_dense_transform(g_flow_h, 2)
    report('RescaleTransform(2)', host(g_flow_f), flow_f)
    # the warp is checked on the ORACLE's field so that its error is its own
    d_flow_f = dev(flow_f, 'planar')
    report('linear warp 160x160x192 (texture-gather kernel)', host(ops.warp(d_img, d_flow_f)), moved)
    os.environ['DFM_WARP_TEX'] = '0'            # the TMA-brick kernel in the regime the round-1 verdict named (bench field)
    try:
        report('linear warp 160x160x192 (TMA-brick kernel)', host(ops.warp(d_img, d_flow_f)), moved)
    finally:
        del os.environ['DFM_WARP_TEX']
    report('linear warp 160x160x192, channels-last field', host(ops.warp(d_img, dev(flow_f, 'cl'))), moved)
    got_nn = host(ops.warp(dev(labels), d_flow_f, 'nearest', 0))
    np.testing.assert_array_equal(got_nn, moved_nn)           # label warps: bit-exact in both builds
    # fused rescale + warp against the two-step oracle (default build: separable evaluation, a few ulp)
    report('fused rescale+warp', host(ops.rescale_warp(d_img, dev(flow_h, 'planar'), 2)), moved)
    # the whole chain end to end (errors of the stages compound; still inside the bar)
    report('chain VecInt -> x2 -> warp', host(ops.warp(d_img, g_flow_f)), moved)


def test_vecint_other_step_counts_at_baseline_shape():
    """config.json int_steps = 5 and the halo-variant boundaries (1, 2, 3 steps) on the bench field."""
    svf = headline_oracle()[0]
    for n in (1, 2, 5):
        report('VecInt(%d) 80x80x96' % n, host(ops.vecint(dev(svf), n)), io.vec_int(svf, n))


@functools.lru_cache(maxsize=None)
def compose_oracle(full):
    svf_a, _ = bench.synth_inputs(1, 'cpu', 3)
    svf_b, _ = bench.synth_inputs(1, 'cpu', 4)
    a, b = io.vec_int(svf_a.numpy(), 7), io.vec_int(svf_b.numpy() / 3.0, 7)       # std-3 and std-1 flows (config 4)
    if full:
        a, b = io.rescale_dense_transform(a, 2), io.rescale_dense_transform(b, 2)
    return a, b, np.stack([io.compose([a[0], b[0]])])


@pytest.mark.parametrize('full', [False, True], ids=['80x80x96', '160x160x192'])
def test_compose_at_baseline_shapes(full):
    a, b, want = compose_oracle(full)
    report('compose %s' % ('full' if full else 'half'), host(ops.compose([dev(a, 'planar'), dev(b, 'planar')])), want)
    report('compose cl', host(ops.compose([dev(a, 'cl'), dev(b, 'cl')])), want)


@functools.lru_cache(maxsize=None)
def c26_oracle():
    """One 160x160x192x26 item (config 3): forward and d/dfield of sum(g * warp) by the torch oracle, in x slabs
    (every output voxel is independent, so slabs of the field reproduce the full autograd result)."""
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(7)
    X, Y, Z, C = 160, 160, 192, 26
    img = torch.rand(X, Y, Z, C, generator=g)
    gout = torch.rand(X, Y, Z, C, generator=g) - 0.5
    flow = torch.from_numpy(headline_oracle()[3][0])          # the bench field at full resolution
    out = torch.empty(X, Y, Z, C)
    gfield = torch.empty(X, Y, Z, 3)
    my, mz = torch.meshgrid(torch.arange(Y, dtype=torch.float32), torch.arange(Z, dtype=torch.float32), indexing='ij')
    S = 8
    for x0 in range(0, X, S):
        f = flow[x0:x0 + S].clone().requires_grad_(True)
        mx = torch.arange(x0, x0 + S, dtype=torch.float32)[:, None, None].expand(S, Y, Z)
        loc = [mx + f[..., 0], my[None] + f[..., 1], mz[None] + f[..., 2]]
        o = to.interpn(img, loc, 'linear')
        (o * gout[x0:x0 + S]).sum().backward()
        out[x0:x0 + S] = o.detach()
        gfield[x0:x0 + S] = f.grad
    return img.numpy(), flow.numpy(), gout.numpy(), out.numpy(), gfield.numpy()


def test_c26_channels_last_forward_and_dfield_at_full_size():
    img, flow, gout, want, want_g = c26_oracle()
    d_img = dev(img[None])                                    # channels-last, the reference layout
    d_flow = dev(flow[None], 'planar').requires_grad_(True)
    out = ops.warp(d_img, d_flow)
    report('C=26 channels-last forward 160x160x192', host(out.detach())[0], want)
    out.backward(dev(gout[None]))
    got_g = host(d_flow.grad)[0]
    # gradients: sums of 26 x 8 products in another order than autograd's; tolerance relative to the gradient scale
    scale = np.abs(want_g).max()
    err = np.abs(got_g.astype(np.float64) - want_g).max()
    print('\n[parity C=26 d/dfield | %s build] max abs err %.3e (gradient scale %.3e)'
          % ('exact' if mrb._lib.exact_order() else 'default', err, scale))
    np.testing.assert_allclose(got_g, want_g, rtol=1e-4, atol=1e-5 * scale)


@functools.lru_cache(maxsize=None)
def jacobian_256_oracle(kind):
    rng = np.random.default_rng(11)
    coarse = torch.from_numpy(rng.standard_normal((1, 3, 16, 16, 16)).astype(np.float32))
    f = torch.nn.functional.interpolate(coarse, size=(256, 256, 256), mode='trilinear', align_corners=True)
    f = f[0].permute(1, 2, 3, 0).contiguous().numpy() * (3.0 if kind == 'smooth' else 12.0)     # few / many folds
    if kind == 'folds':
        f = f + rng.standard_normal(f.shape).astype(np.float32) * 0.3
    det, n_neg = jo.jacobian_determinant(f[:, :, :, None, :].astype(np.float64))     # eval_reg_with_jacobian.py:62-78
    return f, det.reshape(252, 252, 252), int(n_neg)


@pytest.mark.parametrize('kind', ['smooth', 'folds'])
def test_jacobian_256_against_reference_lines(kind):
    f, want, n_neg = jacobian_256_oracle(kind)
    # fp64 field (what get_fdata() hands the reference): all-fp64 kernel
    det, stats = ops.jacobian_determinant(dev(f.astype(np.float64)[None]), out_dtype=torch.float64)
    got = det.cpu().numpy()[0]
    print('\n[parity jacobian 256^3 %s, fp64 field] max abs err %.3e, folds %d (reference %d)'
          % (kind, np.abs(got - want).max(), int(stats[0, 0].item()), n_neg))
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    assert int(stats[0, 0].item()) == n_neg
    assert np.isclose(stats[0, 1].item() / stats[0, 3].item(), want.mean(), rtol=1e-10)
    # fp32 field in both layouts: tiled fp32-stencil kernel (planar) and the direct kernel (channels-last)
    for layout in ('planar', 'cl'):
        det, stats = ops.jacobian_determinant(dev(f[None], layout), out_dtype=torch.float64)
        got = det.cpu().numpy()[0]
        err = np.abs(got - want).max()
        near0 = int((np.abs(want) < 1e-4).sum())
        print('[parity jacobian 256^3 %s, fp32 %s field] max abs err %.3e, folds %d (reference %d, |det| < 1e-4 at %d voxels)'
              % (kind, layout, err, int(stats[0, 0].item()), n_neg, near0))
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)
        # the fold count can only differ at determinants within the fp32 error of zero
        assert abs(int(stats[0, 0].item()) - n_neg) <= near0


# --------------------------------------------------------------------------------------
# plane-marching SS kernel: shapes around its eligibility limits, halo variants, fallback
# --------------------------------------------------------------------------------------
def smooth(rng, shape, std):
    c = rng.standard_normal(shape).astype(np.float32)
    for ax in (1, 2, 3):
        c = (c + np.roll(c, 1, ax) + np.roll(c, -1, ax)) / 3
    return (c / max(c.std(), 1e-6) * std).astype(np.float32)


@pytest.mark.parametrize('shape', [(10, 12, 36), (9, 7, 64), (21, 19, 96), (7, 9, 100), (6, 13, 128), (34, 5, 68), (3, 2, 40)])
@pytest.mark.parametrize('nsteps', [1, 2, 3, 7])
@pytest.mark.parametrize('std', [0.5, 8.0, 60.0])
def test_vecint_marching_kernel(shape, nsteps, std):
    """Z in 33..128 runs k_ss_march: partial z chunks, strips and x segments past the volume edge, both halo
    variants, the per-warp global-gather fallback (std 60: most warps), per-item selection (items differ by 40x)."""
    rng = np.random.default_rng(hash((shape, nsteps)) % 2 ** 31)
    svf = smooth(rng, (3,) + shape + (3,), std)
    svf[1] *= 0.025                                           # per-item bounds differ
    want = io.vec_int(svf, nsteps)

    def check(got):
        if mrb._lib.exact_order():
            np.testing.assert_array_equal(got, want)          # the real check: same bits as the oracle
        elif std <= 1.0:
            np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)
        else:
            # these fields are far rougher than anything in the domain (|grad v| >> 1): squaring them is chaotic and
            # amplifies the fused build's ~1e-7 rounding differences, so the bar holds for all but isolated voxels
            err = np.abs(got.astype(np.float64) - want)
            assert (err > ATOL * std + RTOL * np.abs(want)).mean() < 1e-4 and err.max() < 1e-3 * std

    for layout in ('cl', 'planar'):
        check(host(ops.vecint(dev(svf, layout), nsteps)))
    # differentiable path (save_steps) runs the same kernel on the scaled copy
    s = dev(svf, 'cl').requires_grad_(True)
    check(host(ops.vecint(s, nsteps).detach()))


@pytest.mark.parametrize('shape', [(10, 12, 36), (21, 19, 96), (6, 13, 128), (3, 2, 40)])
@pytest.mark.parametrize('std', [0.5, 8.0, 60.0])
def test_compose_shapes_and_displacements(shape, std):
    """vxm.utils.compose at the shapes the marching SS kernel covers (compose itself runs the bounding-box brick kernel:
    a marching variant with the displacements read from the second field was measured 2x slower on full-size displacements),
    small / large / far-out-of-volume displacements, both layouts."""
    rng = np.random.default_rng(hash(shape) % 2 ** 31)
    a = smooth(rng, (2,) + shape + (3,), std)
    b = smooth(rng, (2,) + shape + (3,), std * 0.5)
    want = np.stack([io.compose([a[i], b[i]]) for i in range(2)])
    for layout in ('cl', 'planar'):
        got = host(ops.compose([dev(a, layout), dev(b, layout)]))
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL * max(1.0, std / 8))
        if mrb._lib.exact_order():
            np.testing.assert_array_equal(got, want)


def test_vecint_marching_nan_and_inf_do_not_hang():
    """NaN / Inf voxels (the reference's int cast of a NaN location is undefined, so there is no oracle for them)
    take the global-gather path without faulting; voxels outside their dependency cone are unaffected."""
    rng = np.random.default_rng(5)
    clean = smooth(rng, (1, 12, 12, 96, 3), 1.0)
    svf = clean.copy()
    svf[0, 6, 6, 40, 1] = np.nan
    svf[0, 2, 3, 70, 0] = np.inf
    want = io.vec_int(clean, 3)
    got = host(ops.vecint(dev(svf), 3))
    torch.cuda.synchronize()
    far = np.ones(svf.shape[1:4], bool)
    far[:, :, 20:] = False                                    # |v| < 4 over 3 steps: z < 20 cannot see z >= 40 - 16
    np.testing.assert_allclose(got[0][far], want[0][far], rtol=RTOL, atol=ATOL)
    assert not np.isfinite(got[0, 6, 6, 40]).all()


# --------------------------------------------------------------------------------------
# randomised sweep (scripts/fuzz_parity.py with a fixed seed budget)
# --------------------------------------------------------------------------------------
def test_fuzz_parity_fixed_seed(arithmetic_mode):
    if arithmetic_mode == 'exact':
        pytest.skip('the sweep switches builds itself; run once')
    sys.path.insert(0, os.path.join(ROOT, 'scripts'))
    import fuzz_parity
    old = sys.argv
    try:
        sys.argv = ['fuzz_parity.py', '12', '20261018']
        assert fuzz_parity.main() == 0
    finally:
        sys.argv = old
        mrb._lib.use(False)
