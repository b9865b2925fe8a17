"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star): nearest-neighbour warps bit-exact; trilinear warps and
integrated fields within 1e-5 relative / 1e-4 voxel absolute; Jacobian determinants within
1e-4.  Every test runs against both builds of the library: the exact-order build keeps the
oracle's op order with separately rounded fp32 ops and is checked for BIT-EXACT equality on the
linear paths; the default fused/packed build is checked against the north_star tolerance (it
differs from the exact build only by removed intermediate roundings).  Gradients (atomics, FMA)
use tolerances.
"""
import glob
import os

import numpy as np
import pytest
import torch

import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
from oracle import interp_oracle as io
from oracle import jacobian_oracle as jo
from oracle import torch_oracle as to

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-4          # north_star tolerance for fp32 linear paths
vxm, ne = mrb.voxelmorph, mrb.neurite


def dev(a, layout='cl'):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return ops.to_layout(t, layout) if t.dim() >= 3 else t


def host(t):
    return ops.to_layout(t, 'cl').cpu().numpy() if t.dim() >= 3 else t.cpu().numpy()


def smooth_noise(rng, shape, std, smooth=2):
    f = rng.standard_normal(shape)
    for _ in range(smooth):
        for ax in range(3):
            f = (np.roll(f, 1, ax + (f.ndim - 4)) + f + np.roll(f, -1, ax + (f.ndim - 4))) / 3.0
    return (f / f.std() * std).astype(np.float32)


@pytest.fixture(autouse=True, params=['fast', 'exact'])
def arithmetic_mode(request):
    """Every test runs against both builds: libdfm.so (fused/packed accumulation, the default
    product) and libdfm_exact.so (reference op order, separately rounded ops)."""
    mrb._lib.use(request.param == 'exact')
    yield request.param
    mrb._lib.use(False)


def assert_linear_parity(got, want):
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)       # north_star bar, both builds
    if mrb._lib.exact_order():
        np.testing.assert_array_equal(got, want)  # exact build: same op order -> same bits


# --------------------------------------------------------------------------------------
# SpatialTransformer forward
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize('shape', [(8, 12, 16), (7, 9, 11), (16, 16, 24)])
@pytest.mark.parametrize('C', [1, 3])
@pytest.mark.parametrize('std', [0.5, 4.0])
@pytest.mark.parametrize('img_layout,field_layout', [('cl', 'cl'), ('planar', 'planar'), ('cl', 'planar')])
def test_warp_linear(shape, C, std, img_layout, field_layout):
    rng = np.random.default_rng(hash((shape, C, std)) % 2 ** 32)
    B = 2
    img = rng.random((B,) + shape + (C,)).astype(np.float32)
    field = smooth_noise(rng, (B,) + shape + (3,), std)
    want = io.spatial_transformer(img, field, 'linear')
    got = host(ops.warp(dev(img, img_layout), dev(field, field_layout), 'linear'))
    assert_linear_parity(got, want)


@pytest.mark.parametrize('shape', [(8, 12, 16), (7, 9, 11)])
@pytest.mark.parametrize('dtype', [np.float32, np.uint8, np.int16, np.int32, np.float64])
@pytest.mark.parametrize('fill', [None, 0])
def test_warp_nearest_bit_exact(shape, dtype, fill):
    rng = np.random.default_rng(11)
    B, C = 2, 2
    img = (rng.random((B,) + shape + (C,)) * 26).astype(dtype)
    field = smooth_noise(rng, (B,) + shape + (3,), 3.0)
    field[0, 0, 0, :4] = [[0.5, 0.5, 0.5], [1.5, -0.5, 2.5], [-3, -3, -3], [50, 50, 50]]   # ties, OOB
    want = io.spatial_transformer(img, field, 'nearest', fill_value=fill)
    for lay in ('cl', 'planar'):
        got = host(ops.warp(dev(img, lay), dev(field, lay), 'nearest', fill_value=fill))
        assert got.dtype == want.dtype
        np.testing.assert_array_equal(got, want)


def test_warp_fill_value_linear_and_mismatched_grid():
    rng = np.random.default_rng(5)
    img = rng.random((1, 10, 12, 8, 2)).astype(np.float32)
    field = smooth_noise(rng, (1, 6, 5, 12, 3), 5.0)       # output grid differs from the image grid
    for fv in (None, -2.5):
        want = io.spatial_transformer(img, field, 'linear', fill_value=fv)
        got = host(ops.warp(dev(img), dev(field), 'linear', fill_value=fv))
        assert got.shape == (1, 6, 5, 12, 2)
        assert_linear_parity(got, want)


@pytest.mark.parametrize('shape', [(1, 12, 16), (8, 1, 16), (8, 12, 1), (1, 1, 5), (2, 2, 4)])
def test_degenerate_axes(shape):
    # size-1 axes (2-D images embedded in 3-D): both corners alias the single voxel
    rng = np.random.default_rng(77)
    img = rng.random((2,) + shape + (2,)).astype(np.float32)
    field = (rng.standard_normal((2,) + shape + (3,)) * 1.5).astype(np.float32)
    for layout in ('cl', 'planar'):
        assert_linear_parity(host(ops.warp(dev(img, layout), dev(field, layout))), io.spatial_transformer(img, field))
        assert_linear_parity(host(ops.warp(dev(img[..., :1], layout), dev(field, layout))),
                             io.spatial_transformer(img[..., :1], field))
        assert_linear_parity(host(ops.vecint(dev(field, layout), 3)), io.vec_int(field, 3))
    np.testing.assert_array_equal(host(ops.warp(dev(img), dev(field), 'nearest')),
                                  io.spatial_transformer(img, field, 'nearest'))
    want = io.rescale_dense_transform(field, 2)
    assert_linear_parity(host(ops.rescale_dense_transform(dev(field, 'planar'), 2)), want)


def test_strong_local_deformation_takes_the_checked_path():
    # white-noise fields: the bounding box of a tile exceeds the brick -> per-thread fallback
    rng = np.random.default_rng(78)
    shape = (16, 16, 64)
    field = (rng.standard_normal((2,) + shape + (3,)) * 6).astype(np.float32)
    img = rng.random((2,) + shape + (1,)).astype(np.float32)
    assert_linear_parity(host(ops.vecint(dev(field, 'planar'), 4)), io.vec_int(field, 4))
    assert_linear_parity(host(ops.warp(dev(img), dev(field, 'planar'))), io.spatial_transformer(img, field))
    assert_linear_parity(host(ops.warp(dev(img), dev(field, 'planar'), fill_value=0.5)),
                         io.spatial_transformer(img, field, fill_value=0.5))


def test_identity_and_translation_known_answers():
    rng = np.random.default_rng(3)
    img = rng.random((1, 8, 8, 8, 1)).astype(np.float32)
    zero = np.zeros((1, 8, 8, 8, 3), np.float32)
    np.testing.assert_array_equal(host(ops.warp(dev(img), dev(zero))), img)
    s = zero.copy()
    s[..., 1] = 3
    got = host(ops.warp(dev(img), dev(s)))
    np.testing.assert_array_equal(got, img[:, :, np.clip(np.arange(8) + 3, 0, 7)])


@pytest.mark.parametrize('shape', [(6, 8, 12, 4), (5, 7, 9, 26), (4, 3, 5, 33), (9, 4, 6, 1)])
def test_transform_channelwise(shape):
    """train_synthmorph.py:67 -- k_warp_cw addresses [X,Y,Z,C] / [X,Y,Z,C,3] in place; the oracle runs the
    reference's 4-D interpn (16 corners, integral channel coordinate)."""
    rng = np.random.default_rng(9)
    X, Y, Z, C = shape
    vol = rng.random((X, Y, Z, C)).astype(np.float32)
    shift = smooth_noise(rng, (X, Y, Z, C, 3), 2.0, smooth=0)
    want = io.transform(vol, shift)
    got = vxm.utils.transform(vol, shift)
    assert tuple(got.shape) == (X, Y, Z, C)
    assert_linear_parity(got.cpu().numpy(), want)
    want_fill = io.transform(vol, shift, fill_value=-1.0)
    assert_linear_parity(vxm.utils.transform(vol, shift, fill_value=-1.0).cpu().numpy(), want_fill)
    # nearest takes the per-channel path
    np.testing.assert_array_equal(vxm.utils.transform(vol, shift, 'nearest').cpu().numpy(), io.transform(vol, shift, 'nearest'))
    # batched call + the fused tf.argmax of train_synthmorph.py:68 (first maximum)
    vol_b = np.stack([vol, vol[::-1].copy()])
    shift_b = np.stack([shift, -shift])
    got_b = ops.warp_channelwise(dev(vol_b), torch.from_numpy(shift_b).cuda())
    assert_linear_parity(got_b[0].cpu().numpy(), want)
    assert_linear_parity(got_b[1].cpu().numpy(), io.transform(vol_b[1], shift_b[1]))
    if C <= 256:
        lab = ops.warp_channelwise(dev(vol_b), torch.from_numpy(shift_b).cuda(), argmax=True)
        assert lab.dtype == torch.uint8 and tuple(lab.shape) == (2, X, Y, Z)
        np.testing.assert_array_equal(lab.cpu().numpy(), got_b.argmax(-1).cpu().numpy().astype(np.uint8))   # same values -> same argmax


def test_channelwise_argmax_ties_go_to_the_first_channel():
    vol = np.zeros((4, 4, 4, 5), np.float32)
    vol[..., 1] = 1.0
    vol[..., 3] = 1.0                                         # channels 1 and 3 tie everywhere
    shift = np.zeros((4, 4, 4, 5, 3), np.float32)
    lab = ops.warp_channelwise(dev(vol[None]), torch.from_numpy(shift[None]).cuda(), argmax=True)
    assert (lab.cpu().numpy() == 1).all()
    lab0 = ops.warp_channelwise(dev(np.zeros_like(vol)[None]), torch.from_numpy(shift[None]).cuda(), argmax=True)
    assert (lab0.cpu().numpy() == 0).all()


def test_interpn_absolute_locations():
    rng = np.random.default_rng(13)
    vol = rng.random((6, 7, 8, 2)).astype(np.float32)
    loc = (rng.random((5, 4, 3)) * 9 - 1).astype(np.float32)
    for method in ('linear', 'nearest'):
        want = io.interpn(vol, loc, method)
        got = ne.utils.interpn(vol, loc, method).cpu().numpy()
        (assert_linear_parity if method == 'linear' else np.testing.assert_array_equal)(got, want)
    want = io.interpn(vol[..., 0], [loc[..., d] for d in range(3)], 'linear', fill_value=0.25)
    got = ne.utils.interpn(vol[..., 0], [loc[..., d] for d in range(3)], 'linear', fill_value=0.25).cpu().numpy()
    assert_linear_parity(got, want)


# --------------------------------------------------------------------------------------
# VecInt / compose / rescale
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize('shape', [(8, 8, 12), (9, 7, 10), (20, 20, 24)])
@pytest.mark.parametrize('nsteps', [0, 1, 2, 5, 7])
@pytest.mark.parametrize('layout', ['cl', 'planar'])
def test_vecint(shape, nsteps, layout):
    rng = np.random.default_rng(hash((shape, nsteps)) % 2 ** 32)
    svf = smooth_noise(rng, (2,) + shape + (3,), 3.0)
    want = io.vec_int(svf, nsteps)
    for out_layout in ('planar', 'cl'):
        got = host(ops.vecint(dev(svf, layout), nsteps, out_layout=out_layout))
        assert_linear_parity(got, want)


@pytest.mark.parametrize('save_steps', [False, True])
def test_vecint_static_halo_per_item_bounds(save_steps):
    # the brick kernel picks its static-halo path per batch item from a bound measured on the device:
    # items whose displacements stay below one voxel for every step, cross the 1- and 2-voxel halos
    # midway, or exceed them from the start must all match the oracle
    rng = np.random.default_rng(123)
    shape = (12, 16, 36)                                     # Z multiple of 4 (TMA path), tiles with edges
    base = smooth_noise(rng, (4,) + shape + (3,), 1.0)
    base /= np.abs(base).max(axis=(1, 2, 3, 4), keepdims=True)
    svf = base * np.array([0.4, 3.0, 9.0, 60.0], np.float32).reshape(4, 1, 1, 1, 1)
    want = io.vec_int(svf, 7)
    t = dev(svf, 'cl')
    if save_steps:
        got = host(ops.vecint(t.requires_grad_(True), 7).detach())
    else:
        got = host(ops.vecint(t, 7))
    assert_linear_parity(got, want)
    if not save_steps:
        # two steps only: after the 2^-2 scaling the last item still moves by more than a voxel, so CTAs of the
        # optimistic channels-last first step fall back to global gathers while the other items stay static
        svf2 = base[:3] * np.array([0.4, 3.0, 9.0], np.float32).reshape(3, 1, 1, 1, 1)
        assert_linear_parity(host(ops.vecint(dev(svf2, 'cl'), 2)), io.vec_int(svf2, 2))


def test_vecint_constant_svf_known_answer():
    c = np.broadcast_to(np.array([1.5, -0.75, 0.25], np.float32), (1, 8, 8, 8, 3)).copy()
    np.testing.assert_array_equal(host(ops.vecint(dev(c), 7)), c)


def test_vecint_layer_and_integrate_vec():
    rng = np.random.default_rng(17)
    svf = smooth_noise(rng, (2, 8, 12, 16, 3), 2.0)
    want = io.vec_int(svf, 5)
    assert_linear_parity(host(vxm.layers.VecInt(int_steps=5)(svf)), want)
    got = vxm.utils.integrate_vec(svf[0], nb_steps=5)
    assert_linear_parity(host(got[None])[0], want[0])


@pytest.mark.parametrize('layout', ['cl', 'planar'])
def test_compose(layout):
    rng = np.random.default_rng(19)
    a = smooth_noise(rng, (2, 8, 12, 16, 3), 3.0)
    b = smooth_noise(rng, (2, 8, 12, 16, 3), 1.0)
    c = smooth_noise(rng, (2, 8, 12, 16, 3), 2.0)
    want = np.stack([io.compose([a[i], b[i]]) for i in range(2)])
    assert_linear_parity(host(ops.compose([dev(a, layout), dev(b, layout)])), want)
    want3 = np.stack([io.compose([a[i], b[i], c[i]]) for i in range(2)])
    assert_linear_parity(host(ops.compose([dev(a, layout), dev(b, layout), dev(c, layout)], out_layout='cl')), want3)
    wantn = np.stack([io.compose([a[i], b[i]], 'nearest') for i in range(2)])
    np.testing.assert_array_equal(host(ops.compose([dev(a, layout), dev(b, layout)], 'nearest')), wantn)
    # the unbatched reference call (bids_two_steps_registration.py:324) + K.eval analogue
    got = vxm.utils.to_numpy(vxm.utils.compose([a[0], b[0]]))
    assert_linear_parity(got, want[0])
    z = np.zeros_like(a)
    np.testing.assert_array_equal(host(ops.compose([dev(a), dev(z)])), a)


@pytest.mark.parametrize('shape,factor', [((8, 8, 12), 2), ((8, 12, 16), 0.5), ((6, 6, 8), 1), ((5, 7, 6), 2),
                                          ((8, 8, 8), 1.5), ((10, 10, 10), 0.3)])
@pytest.mark.parametrize('layout', ['cl', 'planar'])
def test_rescale_dense_transform(shape, factor, layout):
    rng = np.random.default_rng(23)
    f = smooth_noise(rng, (2,) + shape + (3,), 2.0, smooth=1)
    want = io.rescale_dense_transform(f, factor)
    for out_layout in ('planar', 'cl'):
        got = host(ops.rescale_dense_transform(dev(f, layout), factor, out_layout=out_layout))
        assert got.shape == want.shape
        assert_linear_parity(got, want)
    wantn = io.rescale_dense_transform(f, factor, 'nearest')
    np.testing.assert_array_equal(host(ops.rescale_dense_transform(dev(f, layout), factor, 'nearest')), wantn)
    # unbatched + batched dispatch of the reference function (3d_reg.py:394)
    assert_linear_parity(host(vxm.utils.rescale_dense_transform(f, factor)), want)
    assert_linear_parity(host(vxm.utils.rescale_dense_transform(f[0], factor)[None])[0], want[0])


def test_resize_any_channels():
    rng = np.random.default_rng(29)
    vol = rng.random((6, 8, 8, 5)).astype(np.float32)
    want = io.resize(vol, [2, 1.5, 0.5])
    got = host(ne.utils.resize(vol, [2, 1.5, 0.5])[None])[0]
    assert_linear_parity(got, want)


def test_transform_model_and_vxmdense_tail():
    rng = np.random.default_rng(31)
    scan = rng.random((2, 8, 12, 16, 1)).astype(np.float32)
    half = smooth_noise(rng, (2, 4, 6, 8, 3), 1.5, smooth=1)
    for interp in ('linear', 'nearest'):
        want = io.transform_model(scan, half, interp, rescale=2)
        got = vxm.networks.Transform((8, 12, 16), interp_method=interp, rescale=2).predict([scan, half])
        (assert_linear_parity if interp == 'linear' else np.testing.assert_array_equal)(got, want)
    full = smooth_noise(rng, (2, 8, 12, 16, 3), 2.0)
    want = io.transform_model(scan, full, 'linear', rescale=1)       # 3d_reg.py:317,333: scale == 1
    got = vxm.networks.Transform((8, 12, 16), rescale=1).predict([scan, full])
    assert_linear_parity(got, want)
    # VxmDense tail: svf at half res -> VecInt(5) -> x2 -> linear warp; second output = preint flow
    fused = vxm.networks.VxmDense((8, 12, 16), int_steps=5, svf_resolution=2, int_resolution=2, fuse_rescale_warp=True)
    model = vxm.networks.VxmDense((8, 12, 16), int_steps=5, svf_resolution=2, int_resolution=2)
    # fused = unfused bit for bit in the exact build; the fast build's up-sampler is separable (few ulp)
    assert_linear_parity(fused.predict_deform([scan, half])[0], model.predict_deform([scan, half])[0])
    assert fused.references.pos_flow is None
    y, pre = model.predict_deform([scan, half])
    model.deform([scan, half])          # keeps references.pos_flow (predict_deform may fuse it away)
    flow = io.rescale_dense_transform(io.vec_int(half, 5), 2)
    assert_linear_parity(y, io.spatial_transformer(scan, flow))
    np.testing.assert_array_equal(pre, half)
    assert_linear_parity(host(model.references.pos_flow), flow)


def test_chunked_predict_equals_single_call():
    rng = np.random.default_rng(67)
    scan = rng.random((5, 8, 12, 16, 1)).astype(np.float32)
    half = smooth_noise(rng, (5, 4, 6, 8, 3), 1.5, smooth=1)
    model = vxm.networks.VxmDense((8, 12, 16), int_steps=5, svf_resolution=2, int_resolution=2)
    y1, p1 = model.predict_deform([scan, half], batch_size=8)          # one call
    y2, p2 = model.predict_deform([scan, half], batch_size=2)          # 3 chunks on side streams
    np.testing.assert_array_equal(y1, y2)
    np.testing.assert_array_equal(p1, half)
    np.testing.assert_array_equal(p2, half)
    warp = vxm.networks.VxmDense((8, 12, 16), int_steps=5, svf_resolution=2, int_resolution=2, reg_field='warp')
    y3, w3 = warp.predict_deform([scan, half], batch_size=2)
    np.testing.assert_array_equal(y3, y1)
    assert_linear_parity(w3, io.rescale_dense_transform(io.vec_int(half, 5), 2))


def test_stitch_subvolumes_against_reference_goldens():
    import importlib.util
    from oracle import stitch_oracle as so
    here = os.path.dirname(__file__)
    spec = importlib.util.spec_from_file_location('make_stitch_golden', os.path.join(here, 'golden', 'make_stitch_golden.py'))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    paths = sorted(glob.glob(os.path.join(here, 'golden', 'stitch_*.npz')))
    assert len(paths) >= 3
    for path in paths:
        g = np.load(path)
        coords = [tuple(int(v) for v in c) for c in g['coords']]
        warps = gen.make_warps(int(g['seed']), g['in_shape'], len(coords))
        got = ops.stitch_subvolumes(tuple(g['in_shape']), tuple(g['vol_shape']), coords, warps)
        assert got.dtype == torch.float64
        np.testing.assert_array_equal(got.cpu().numpy(), g['out'])          # bit-identical to the reference
        got32 = ops.stitch_subvolumes(tuple(g['in_shape']), tuple(g['vol_shape']), coords,
                                      ops.to_layout(torch.from_numpy(np.stack(warps)).cuda(), 'planar'), out_dtype=torch.float32)
        np.testing.assert_allclose(got32.cpu().numpy(), g['out'], rtol=1e-6, atol=1e-6)
    # uncovered voxels stay zero
    f = np.random.default_rng(0).standard_normal((8, 8, 8, 3)).astype(np.float32)
    out = ops.stitch_subvolumes((8, 8, 8), (12, 8, 8), [(0, 8, 0, 8, 0, 8)], [f]).cpu().numpy()
    np.testing.assert_array_equal(out, so.get_def_field_from_subvol((8, 8, 8), (12, 8, 8), [(0, 8, 0, 8, 0, 8)], [f]))
    with pytest.raises(ValueError):
        ops.stitch_subvolumes((8, 8, 8), (12, 8, 8), [(6, 14, 0, 8, 0, 8)], [f])


def test_cuda_graph_replay_matches_eager():
    rng = np.random.default_rng(73)
    model = vxm.networks.VxmDense((16, 16, 32), int_steps=7, svf_resolution=2, int_resolution=2)
    for trial in range(3):                                   # same shapes -> one capture, three replays
        scan = rng.random((1, 16, 16, 32, 1)).astype(np.float32)
        half = smooth_noise(rng, (1, 8, 8, 16, 3), 1.5, smooth=1)
        eager = [host(t) if t.dim() == 5 else t for t in model.deform([dev(scan), dev(half)])]
        y, pre = model.deform_graphed([dev(scan), dev(half)])
        np.testing.assert_array_equal(host(y), eager[0])
        np.testing.assert_array_equal(host(pre), half)
    assert len(model._graphs) == 1


def test_fused_rescale_warp_matches_unfused_bitwise():
    rng = np.random.default_rng(61)
    for shape, B in [((8, 12, 16), 2), ((20, 20, 48), 1), ((6, 5, 7), 1)]:
        full = tuple(2 * s for s in shape)
        half = smooth_noise(rng, (B,) + shape + (3,), 2.5, smooth=1)
        scan = rng.random((B,) + full + (1,)).astype(np.float32)
        for fv in (None, 0.0):
            fused = host(ops.rescale_warp(dev(scan), dev(half, 'planar'), 2, fv))
            flow = ops.rescale_dense_transform(dev(half, 'planar'), 2)
            unfused = host(ops.warp(dev(scan), flow, 'linear', fv))
            if mrb._lib.exact_order():
                np.testing.assert_array_equal(fused, unfused)       # same arithmetic
            elif fv is None:
                assert_linear_parity(fused, unfused)                # fast build: the stand-alone up-sampler is separable
            else:                                                   # a few-ulp flow change can flip a voxel across the fill boundary
                assert np.mean(~np.isclose(fused, unfused, rtol=RTOL, atol=ATOL)) < 1e-3
            assert_linear_parity(fused, io.spatial_transformer(scan, io.rescale_dense_transform(half, 2), 'linear', fv))


def test_warp_onehot_equals_generic_warp_of_the_onehot_tensor():
    """dfm_warp_onehot_fwd / _bwd (pred of train_synthmorph.py:298 from the label map): the forward pass carries the same bits
    as the generic channels-last warp of the materialised one-hot tensor (both builds), the field gradient agrees with the
    generic adjoint and with fp64 autograd of the oracle; fill values, channels-last fields, label volume larger than the grid."""
    rng = np.random.default_rng(615)
    for shape, B, C, fv, lshape, lay in [((8, 12, 16), 2, 26, None, None, 'planar'), ((9, 7, 13), 1, 5, 0.0, None, 'cl'),
                                         ((8, 12, 16), 3, 26, -1.0, (11, 12, 20), 'planar'), ((20, 20, 48), 1, 40, None, None, 'cl')]:
        flow = smooth_noise(rng, (B,) + shape + (3,), 2.5, smooth=1)
        labels = rng.integers(0, C, (B,) + (lshape or shape)).astype(np.int64)
        onehot = np.eye(C, dtype=np.float32)[labels]
        d_lab = torch.from_numpy(labels).cuda()
        d_flow = dev(flow, lay).requires_grad_(True)
        got = ops.warp_onehot(d_lab, d_flow, C, fv)
        d_flow2 = dev(flow, lay).requires_grad_(True)
        ref = ops.warp(dev(onehot), d_flow2, 'linear', fv)
        np.testing.assert_array_equal(host(got.detach()), host(ref.detach()))
        assert_linear_parity(host(got.detach()), io.spatial_transformer(onehot, flow, 'linear', fv))
        g = torch.from_numpy(rng.standard_normal(tuple(got.shape)).astype(np.float32)).cuda()
        got.backward(g)
        ref.backward(g)
        np.testing.assert_allclose(host(d_flow.grad), host(d_flow2.grad), rtol=1e-4, atol=1e-4 * float(np.abs(host(d_flow2.grad)).max()))
    # more labels than the shared output tile holds: the per-element kernel (forward only; gradients take the generic path)
    lab60 = rng.integers(0, 60, (1, 8, 12, 16)).astype(np.int64)
    flow60 = dev(smooth_noise(rng, (1, 8, 12, 16, 3), 2.0, smooth=1))
    with torch.no_grad():
        assert torch.equal(ops.warp_onehot(torch.from_numpy(lab60).cuda(), flow60, 60, 0.5),
                           ops.to_layout(ops.warp(dev(np.eye(60, dtype=np.float32)[lab60]), flow60, 'linear', 0.5), 'cl'))
    f60 = flow60.clone().requires_grad_(True)
    ops.warp_onehot(torch.from_numpy(lab60).cuda(), f60, 60).sum().backward()
    assert torch.isfinite(f60.grad).all()
    # no-grad path and label dtypes
    lab = rng.integers(0, 26, (2, 8, 12, 16, 1)).astype(np.float32)
    flow = dev(smooth_noise(rng, (2, 8, 12, 16, 3), 2.0, smooth=1))
    with torch.no_grad():
        a = ops.warp_onehot(dev(lab), flow, 26)
        b = ops.warp(dev(np.eye(26, dtype=np.float32)[lab[..., 0].astype(np.int64)]), flow)
    assert torch.equal(a, ops.to_layout(b, 'cl'))


class warp_kernel:
    """Select the one-channel linear warp kernel for the calls inside: 'tex' (texture gathers, the default where the
    image qualifies) or 'brick' (TMA bounding-box brick); the library reads DFM_WARP_TEX per call."""

    def __init__(self, which):
        self.value = '1' if which == 'tex' else '0'

    def __enter__(self):
        self.old = os.environ.get('DFM_WARP_TEX')
        os.environ['DFM_WARP_TEX'] = self.value

    def __exit__(self, *exc):
        if self.old is None:
            del os.environ['DFM_WARP_TEX']
        else:
            os.environ['DFM_WARP_TEX'] = self.old


def test_linear_warp_kernels_agree_bitwise():
    """The texture-gather warp and the TMA-brick warp share weights, corner order and accumulation: same bits, and both
    inside the bar against the oracle (planar and channels-last fields, fill values, image larger than the grid, B > 32)."""
    rng = np.random.default_rng(613)
    cases = [((16, 24, 32), 2, None, None, 'planar'), ((16, 24, 32), 34, None, None, 'cl'), ((18, 22, 40), 1, 0.0, None, 'planar'),
             ((16, 24, 32), 2, -1.0, (20, 28, 48), 'cl'), ((40, 40, 96), 2, None, None, 'planar')]
    for shape, B, fv, img_shape, lay in cases:
        flow = smooth_noise(rng, (B,) + shape + (3,), 3.0, smooth=1)
        scan = rng.random((B,) + (img_shape or shape) + (1,)).astype(np.float32)
        d_scan, d_flow = dev(scan), dev(flow, lay)
        with warp_kernel('tex'):
            a = host(ops.warp(d_scan, d_flow, 'linear', fv))
        with warp_kernel('brick'):
            b = host(ops.warp(d_scan, d_flow, 'linear', fv))
        np.testing.assert_array_equal(a, b)
        if B <= 2:
            assert_linear_parity(a, io.spatial_transformer(scan, flow, 'linear', fv))


def test_fused_texture_gather_path():
    """dfm_rescale_warp_fwd on shapes its texture-gather kernel covers (dfm_warp_tex.cu): same arithmetic as the marching
    up-sampler followed by the warp, so it must equal the two stand-alone kernels BIT FOR BIT in both builds, and the
    oracle within the bar.  Cases: more than one launch's worth of texture objects (B > 32), ragged tiles (Y, Z not
    multiples of 16 / 32), a non-integer zoom, an image larger than the field grid, fill values."""
    rng = np.random.default_rng(611)
    cases = [((8, 12, 16), 33, 2, None, None), ((9, 11, 20), 1, 2, 0.0, None), ((8, 12, 16), 1, 1.5, None, None),
             ((8, 12, 16), 2, 2, -1.0, (20, 24, 32)), ((20, 24, 40), 3, 2, None, None)]
    for shape, B, factor, fv, img_shape in cases:
        full = tuple(int(s * factor) for s in shape)
        half = smooth_noise(rng, (B,) + shape + (3,), 2.5, smooth=1)
        scan = rng.random((B,) + (img_shape or full) + (1,)).astype(np.float32)
        d_scan, d_half = dev(scan), dev(half, 'planar')
        fused = host(ops.rescale_warp(d_scan, d_half, factor, fv))
        flow = ops.rescale_dense_transform(d_half, factor)
        unfused = host(ops.warp(d_scan, flow, 'linear', fv))
        np.testing.assert_array_equal(fused, unfused)
        if B <= 3:
            want = io.spatial_transformer(scan, io.rescale_dense_transform(half, factor), 'linear', fv)
            if fv is None or mrb._lib.exact_order():
                assert_linear_parity(fused, want)
            else:                                                   # a few-ulp flow change can flip a voxel across the fill boundary
                assert np.mean(~np.isclose(fused, want, rtol=RTOL, atol=ATOL)) < 1e-3


def test_fused_rescale_nearest_warp():
    """dfm_rescale_warp_nearest_fwd (label maps through Transform(interp_method='nearest', rescale=2), 3d_reg.py:377-380):
    the same bits as the two stand-alone kernels and as the oracle, float32 and int32 labels, fill values, shapes the
    marching kernel covers and one it does not (work-buffer fallback)."""
    rng = np.random.default_rng(614)
    for shape, B, factor, fv, dt in [((8, 12, 16), 2, 2, None, np.float32), ((9, 11, 20), 1, 2, 0, np.float32),
                                     ((8, 12, 16), 33, 2, 7, np.int32), ((6, 5, 7), 1, 2, 0, np.float32),
                                     ((20, 24, 40), 2, 2, None, np.int32)]:
        full = tuple(int(s * factor) for s in shape)
        half = smooth_noise(rng, (B,) + shape + (3,), 2.5, smooth=1)
        labels = rng.integers(0, 26, (B,) + full + (1,)).astype(dt)
        d_lab, d_half = dev(labels), dev(half, 'planar')
        fused = host(ops.rescale_warp(d_lab, d_half, factor, fv, 'nearest'))
        assert fused.dtype == dt
        unfused = host(ops.warp(d_lab, ops.rescale_dense_transform(d_half, factor), 'nearest', fv))
        np.testing.assert_array_equal(fused, unfused)
        if B <= 2 and mrb._lib.exact_order():          # the default build's up-sampler is a few ulp off: ties may round the other way
            np.testing.assert_array_equal(fused, io.spatial_transformer(labels, io.rescale_dense_transform(half, factor), 'nearest', fv))
    # through the model mirror
    scan = rng.integers(0, 5, (2, 16, 24, 32, 1)).astype(np.float32)
    half = smooth_noise(rng, (2, 8, 12, 16, 3), 2.0, smooth=1)
    got = vxm.networks.Transform((16, 24, 32), interp_method='nearest', rescale=2).predict([scan, half])
    want = host(ops.warp(dev(scan), ops.rescale_dense_transform(dev(half, 'planar'), 2), 'nearest'))
    np.testing.assert_array_equal(got, want)


def test_texture_descriptor_cache_eviction():
    """The host-side cache of texture descriptors with a small capacity (own process: the capacity is read when the library
    loads): hundreds of distinct image buffers cycle through it, every result stays bit-equal to the TMA-brick kernel."""
    import subprocess, sys, textwrap
    code = textwrap.dedent('''
        import os, sys
        import numpy as np, torch
        sys.path.insert(0, %r)
        import multimodal_registration_b200 as mrb
        from multimodal_registration_b200 import ops
        rng = np.random.default_rng(5)
        flow = torch.from_numpy((rng.standard_normal((3, 16, 24, 32, 3)) * 2).astype(np.float32)).cuda()
        half = ops.to_layout(torch.from_numpy((rng.standard_normal((3, 8, 12, 16, 3)) * 1.5).astype(np.float32)).cuda(), 'planar')
        keep, bad = [], 0   # (planar: the stand-alone up-sampler then runs the same marching arithmetic as the fused kernel)
        for it in range(120):                               # 360 descriptors through a cache of 64 (+ 64 in the graveyard)
            img = torch.rand((3, 16, 24, 32, 1), device='cuda')
            keep.append(img)                                # distinct addresses: nothing is freed
            a = ops.warp(img, flow)
            f = ops.rescale_warp(img, half, 2)
            os.environ['DFM_WARP_TEX'] = '0'
            b = ops.warp(img, flow)
            g = ops.warp(img, ops.rescale_dense_transform(half, 2))
            del os.environ['DFM_WARP_TEX']
            bad += int(not torch.equal(a, b)) + int(not torch.equal(f, g))
        # the first buffers' descriptors are long gone: they are re-created on demand
        bad += int(not torch.equal(ops.warp(keep[0], flow), ops.warp(keep[0].clone(), flow)))
        torch.cuda.synchronize()
        print('BAD', bad)
    ''') % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DFM_TEX_CACHE_CAP='64')
    if mrb._lib.exact_order():
        env['DFM_EXACT'] = '1'
    res = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    assert 'BAD 0' in res.stdout, res.stdout[-500:]


def test_texture_kernels_inside_cuda_graph_capture():
    """Texture descriptors are created on first use; the first use may be inside a stream capture (descriptor creation is not a
    stream operation, the library relaxes the capture mode around it).  Replays read the current buffer contents."""
    rng = np.random.default_rng(616)
    half = dev(smooth_noise(rng, (2, 8, 12, 16, 3), 2.0, smooth=1), 'planar')
    scan = torch.rand((2, 16, 24, 32, 1), device='cuda')             # a fresh buffer: no descriptor cached for it yet
    flow = ops.rescale_dense_transform(half, 2)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fused = ops.rescale_warp(scan, half, 2)
        plain = ops.warp(scan, flow)
    for _ in range(2):
        scan.copy_(torch.rand_like(scan))                           # new contents, same address
        g.replay()
        torch.cuda.synchronize()
        with warp_kernel('brick'):
            want = ops.warp(scan, flow)
        assert torch.equal(plain, want) and torch.equal(fused, want)


def test_fused_texture_gather_is_reproducible():
    """The coarse-plane ring of the fused kernel is released by data-dependent arrivals (a consumer's arrival must not
    overtake its shared loads): many launches at a size with thousands of CTAs give the same bits every time."""
    rng = np.random.default_rng(612)
    half = dev(smooth_noise(rng, (4, 40, 40, 48, 3), 3.0, smooth=1), 'planar')
    scan = dev(rng.random((4, 80, 80, 96, 1)).astype(np.float32))
    ref = ops.warp(scan, ops.rescale_dense_transform(half, 2))
    for _ in range(25):
        assert torch.equal(ops.rescale_warp(scan, half, 2), ref)


# --------------------------------------------------------------------------------------
# Jacobian determinant
# --------------------------------------------------------------------------------------
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'jacobian_*.npz')))


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_jacobian_against_reference_golden(path):
    g = np.load(path)
    field = g['field'][None]
    for layout in ('cl', 'planar'):
        for in_dtype in (np.float32, np.float64):
            det, stats = ops.jacobian_determinant(dev(field.astype(in_dtype), layout))
            det = det.cpu().numpy().reshape(-1)
            np.testing.assert_allclose(det, g['det'], rtol=0, atol=1e-4)       # north_star bar
            # fp64 inputs and channels-last fields take the all-fp64 kernel; planar fp32 fields the
            # plane-marching kernel whose stencil differences are formed in fp32
            tight = 1e-12 if (in_dtype == np.float64 or layout == 'cl') else 1e-5
            np.testing.assert_allclose(det, g['det'], rtol=tight, atol=tight)
            n_neg, s, s2, n = stats.cpu().numpy()[0]
            assert int(n_neg) == int(g['n_neg']) and int(n) == g['det'].size
            assert abs(s / n - float(g['mean'])) < max(tight, 1e-12)
            assert abs(np.sqrt(max(s2 / n - (s / n) ** 2, 0)) - float(g['std'])) < max(tight, 1e-9)
    det32, _ = ops.jacobian_determinant(dev(field), out_dtype=torch.float32, want_stats=False)
    np.testing.assert_allclose(det32.cpu().numpy().reshape(-1), g['det'], rtol=0, atol=1e-4)


def test_jacobian_batched_and_oracle():
    rng = np.random.default_rng(37)
    f = smooth_noise(rng, (3, 40, 23, 37, 3), 1.0)      # several tiles with ragged edges
    det, stats = ops.jacobian_determinant(dev(f, 'planar'))
    for b in range(3):
        d, n = jo.jacobian_determinant(f[b][:, :, :, None, :])
        np.testing.assert_allclose(det[b].cpu().numpy().reshape(-1), d, rtol=1e-5, atol=1e-5)
        assert int(stats[b, 0].item()) == n
    det64, _ = ops.jacobian_determinant(dev(f.astype(np.float64), 'planar'))
    for b in range(3):
        d, n = jo.jacobian_determinant(f[b][:, :, :, None, :])
        np.testing.assert_allclose(det64[b].cpu().numpy().reshape(-1), d, rtol=1e-12, atol=1e-12)


# --------------------------------------------------------------------------------------
# backward passes (oracle = autograd of the torch restatement; fp32 kernels vs fp64 oracle)
# --------------------------------------------------------------------------------------
def _grads_oracle(fn, *arrs):
    ts = [torch.from_numpy(a.astype(np.float64)).requires_grad_(True) for a in arrs]
    out = fn(*ts)
    g = torch.from_numpy(np.random.default_rng(1).standard_normal(tuple(out.shape))).to(out.dtype)
    out.backward(g)
    return g.numpy().astype(np.float32), [t.grad.numpy() for t in ts]


@pytest.mark.parametrize('C,layout', [(1, 'cl'), (3, 'cl'), (3, 'planar')])
@pytest.mark.parametrize('fill', [None, 0.0])
def test_warp_backward(C, layout, fill):
    rng = np.random.default_rng(41)
    img = rng.random((2, 6, 8, 10, C))
    field = smooth_noise(rng, (2, 6, 8, 10, 3), 2.5).astype(np.float64)
    g, (gi, gf) = _grads_oracle(lambda i, f: to.spatial_transformer(i, f, 'linear', fill), img, field)
    ti = dev(img.astype(np.float32), layout).requires_grad_(True)
    tf_ = dev(field.astype(np.float32), layout).requires_grad_(True)
    out = ops.warp(ti, tf_, 'linear', fill)
    out.backward(dev(g, 'cl'))
    np.testing.assert_allclose(host(ti.grad), gi, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(host(tf_.grad), gf, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize('shape,C,std', [((6, 8, 10), 5, 2.5), ((12, 20, 40), 7, 3.0), ((9, 9, 36), 3, 9.0)])
@pytest.mark.parametrize('fill', [None, 0.0])
def test_warp_backward_field_only_planar_multichannel(shape, C, std, fill):
    # the `pred` gradient of the training step: planar C-channel image without gradient, d/dfield only
    rng = np.random.default_rng(71)
    img = rng.random((2,) + shape + (C,))
    field = smooth_noise(rng, (2,) + shape + (3,), std).astype(np.float64)
    g, (_, gf) = _grads_oracle(lambda i, f: to.spatial_transformer(i, f, 'linear', fill), img, field)
    ti = dev(img.astype(np.float32), 'planar')
    tf_ = dev(field.astype(np.float32), 'planar').requires_grad_(True)
    out = ops.warp(ti, tf_, 'linear', fill)
    assert_linear_parity(host(out.detach()), io.spatial_transformer(img.astype(np.float32), field.astype(np.float32), 'linear', fill))
    out.backward(dev(g, 'planar'))
    np.testing.assert_allclose(host(tf_.grad), gf, rtol=2e-4, atol=5e-5)


@pytest.mark.parametrize('C', [2, 3, 5, 8, 13, 17, 26, 40])
@pytest.mark.parametrize('field_layout,fill', [('cl', None), ('planar', 0.0)])
def test_warp_channels_last_multichannel_fwd_bwd(C, field_layout, fill):
    # the reference layout (lanes-over-channels kernel): every lane mapping (2..32 lanes per voxel, 32-channel
    # chunks), a voxel count that is not a multiple of 32, forward bit-parity and both gradients
    rng = np.random.default_rng(1000 + C)
    shape = (5, 7, 9)
    img = rng.random((2,) + shape + (C,))
    field = smooth_noise(rng, (2,) + shape + (3,), 2.5).astype(np.float64)
    g, (gi, gf) = _grads_oracle(lambda i, f: to.spatial_transformer(i, f, 'linear', fill), img, field)
    ti = dev(img.astype(np.float32), 'cl').requires_grad_(True)
    tf_ = dev(field.astype(np.float32), field_layout).requires_grad_(True)
    out = ops.warp(ti, tf_, 'linear', fill)
    assert ops.layout_of(out) == 'cl'
    assert_linear_parity(host(out.detach()), io.spatial_transformer(img.astype(np.float32), field.astype(np.float32), 'linear', fill))
    out.backward(dev(g, 'cl'))
    np.testing.assert_allclose(host(ti.grad), gi, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(host(tf_.grad), gf, rtol=2e-4, atol=5e-5)
    # d/dfield only (image without gradient): the training step's `pred` gradient
    tf2 = dev(field.astype(np.float32), field_layout).requires_grad_(True)
    ops.warp(dev(img.astype(np.float32), 'cl'), tf2, 'linear', fill).backward(dev(g, 'cl'))
    np.testing.assert_allclose(host(tf2.grad), gf, rtol=2e-4, atol=5e-5)


@pytest.mark.parametrize('nsteps', [1, 3, 5])
def test_vecint_backward(nsteps):
    rng = np.random.default_rng(43)
    svf = smooth_noise(rng, (2, 6, 8, 10, 3), 2.0).astype(np.float64)
    g, (gs,) = _grads_oracle(lambda s: to.vec_int(s, nsteps), svf)
    ts = dev(svf.astype(np.float32)).requires_grad_(True)
    ops.vecint(ts, nsteps).backward(dev(g, 'cl'))
    np.testing.assert_allclose(host(ts.grad), gs, rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize('shape', [(6, 8, 10), (19, 5, 33), (20, 12, 96), (6, 5, 130)])
@pytest.mark.parametrize('std', [0.3, 6.0, 40.0])
def test_vecint_backward_gather_and_scatter_items(shape, std):
    """Per-item selection between the atomics-free gather adjoint (|v| < 1: early steps) and the scatter adjoint:
    item 1 moves 40x less than item 0, so the two take different kernels in the same launch; segment / strip / z edges
    (shapes that are no multiples of the tile), Z > 128 (scatter only)."""
    rng = np.random.default_rng(shape[0] * 7 + int(std))
    svf = smooth_noise(rng, (2,) + shape + (3,), std).astype(np.float64)
    svf[1] *= 0.025
    nsteps = 5
    g, (gs,) = _grads_oracle(lambda s: to.vec_int(s, nsteps), svf)
    ts = dev(svf.astype(np.float32)).requires_grad_(True)
    ops.vecint(ts, nsteps).backward(dev(g, 'cl'))
    scale = max(1.0, float(np.abs(gs).max()))
    # a sample that lands within an ulp of a cell boundary picks the other cell in fp32 than in the fp64 oracle, and the
    # gradient is discontinuous there: tolerate a handful of such voxels
    bad = ~np.isclose(host(ts.grad), gs, rtol=5e-4, atol=5e-5 * scale)
    assert bad.sum() <= max(3, 2e-3 * bad.size), (bad.sum(), bad.size)


def test_ss_step_bwd_gather_equals_scatter():
    """dfm_ss_step_bwd_bounded with a bound below one voxel (gather kernel, no atomics) against the same call without
    a bound (scatter kernel): same gradient up to summation order; the gather result is bit-reproducible."""
    from multimodal_registration_b200.ops import _ptr, _stream
    rng = np.random.default_rng(77)
    B, X, Y, Z = 3, 21, 14, 40
    v = (smooth_noise(rng, (B, X, Y, Z, 3), 0.18)).astype(np.float32)
    v[2] *= 9.0                                                 # item 2 exceeds the bound -> scatter path
    tv = dev(v, 'planar')
    tg = dev(rng.standard_normal((B, X, Y, Z, 3)).astype(np.float32), 'planar')
    bound = torch.tensor([float(np.abs(v[b]).max()) for b in range(B)], device='cuda')
    assert bound[0] < 1 and bound[1] < 1 and bound[2] > 1
    outs = []
    for use_bound in (False, True, True):
        gv = ops.empty((B, X, Y, Z, 3), 'planar', tv.device)
        gv.fill_(float('nan'))                                  # the call must overwrite / zero everything itself
        if use_bound:
            mrb._lib.call('dfm_ss_step_bwd_bounded', _ptr(tg), _ptr(tv), _ptr(gv), _ptr(bound), 1.0, B, X, Y, Z, 0.5, _stream())
        else:
            mrb._lib.call('dfm_ss_step_bwd', _ptr(tg), _ptr(tv), _ptr(gv), B, X, Y, Z, 0.5, _stream())
        outs.append(host(gv))
    np.testing.assert_allclose(outs[1], outs[0], rtol=2e-5, atol=2e-5)
    np.testing.assert_array_equal(outs[1][:2], outs[2][:2])     # gather items: deterministic


@pytest.mark.parametrize('factor', [2, 0.5, 1.5])
def test_rescale_backward_and_adjointness(factor):
    rng = np.random.default_rng(47)
    f = smooth_noise(rng, (2, 6, 8, 8, 3), 1.0).astype(np.float64)
    g, (gf,) = _grads_oracle(lambda t: to.rescale_dense_transform(t, factor), f)
    tf_ = dev(f.astype(np.float32)).requires_grad_(True)
    out = ops.rescale_dense_transform(tf_, factor)
    out.backward(dev(g, 'cl'))
    np.testing.assert_allclose(host(tf_.grad), gf, rtol=1e-4, atol=1e-5)
    # <K x, y> == <x, K^T y>
    lhs = float((host(out.detach()).astype(np.float64) * g).sum())
    rhs = float((f * host(tf_.grad)).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


@pytest.mark.parametrize('shape,factor', [((20, 12, 24), 2), ((7, 9, 33), 2), ((10, 6, 8), 3), ((16, 16, 16), 1.5)])
def test_rescale_backward_separable_equals_one_pass(shape, factor):
    """dfm_resize_bwd_ws with a workspace (two separable passes: x,y then z) against dfm_resize_bwd (one pass, K^3 gathers)."""
    from multimodal_registration_b200 import _coords
    from multimodal_registration_b200.ops import _ptr, _stream
    rng = np.random.default_rng(5)
    B, C = 2, 3
    Xi, Yi, Zi = shape
    Xo, Yo, Zo = (int(d * factor) for d in shape)
    gout = dev(rng.standard_normal((B, Xo, Yo, Zo, C)).astype(np.float32), 'planar')
    t = [_coords.device_adjoint_taps(a, b, 0) for a, b in ((Xi, Xo), (Yi, Yo), (Zi, Zo))]
    assert max(tt[3] for tt in t) >= 3
    outs = []
    for ws in (False, True):
        gin = ops.empty((B, Xi, Yi, Zi, C), 'planar', gout.device)
        gin.fill_(float('nan'))
        work = torch.empty(mrb._lib.load().dfm_resize_bwd_workspace_bytes(B, C, Xi, Yi, Zo) // 4, device='cuda') if ws else None
        mrb._lib.call('dfm_resize_bwd_ws', _ptr(gout), _ptr(gin), _ptr(work),
                      _ptr(t[0][0]), _ptr(t[0][1]), _ptr(t[0][2]), t[0][3], _ptr(t[1][0]), _ptr(t[1][1]), _ptr(t[1][2]), t[1][3],
                      _ptr(t[2][0]), _ptr(t[2][1]), _ptr(t[2][2]), t[2][3], B, C, Xi, Yi, Zi, Xo, Yo, Zo, float(factor), 1.0, _stream())
        outs.append(host(gin))
    np.testing.assert_allclose(outs[1], outs[0], rtol=2e-5, atol=2e-5)


def test_jacobian_exact_switch_matches_float64_reference():
    """ops.jacobian_determinant(exact=True): a planar fp32 field evaluated in float64 like the reference lines."""
    rng = np.random.default_rng(8)
    f = smooth_noise(rng, (1, 14, 12, 16, 3), 1.5, smooth=1)
    det, stats = ops.jacobian_determinant(dev(f, 'planar'), exact=True)
    d, n = jo.jacobian_determinant(f[0][:, :, :, None, :].astype(np.float64))
    np.testing.assert_allclose(det.cpu().numpy().reshape(-1), d, rtol=0, atol=1e-12)
    assert int(stats[0, 0]) == n


def test_compose_backward():
    rng = np.random.default_rng(53)
    a = smooth_noise(rng, (1, 6, 8, 10, 3), 2.0).astype(np.float64)
    b = smooth_noise(rng, (1, 6, 8, 10, 3), 1.0).astype(np.float64)
    g, (ga, gb) = _grads_oracle(lambda x, y: to.compose([x[0], y[0]])[None], a, b)
    ta = dev(a.astype(np.float32)).requires_grad_(True)
    tb = dev(b.astype(np.float32)).requires_grad_(True)
    ops.compose([ta, tb]).backward(dev(g, 'cl'))
    np.testing.assert_allclose(host(ta.grad), ga, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(host(tb.grad), gb, rtol=1e-4, atol=2e-5)


def test_training_tail_gradient_chain():
    """config.json-shaped chain at toy size: svf -> VecInt(5) -> x2 -> linear warp of a C-channel
    one-hot map, gradient w.r.t. the svf (train_synthmorph.py:296-298,305-306)."""
    rng = np.random.default_rng(59)
    svf = smooth_noise(rng, (1, 4, 6, 8, 3), 1.0).astype(np.float64)
    lab = rng.integers(0, 4, (1, 8, 12, 16))
    onehot = np.eye(4)[lab]
    g, (gs,) = _grads_oracle(
        lambda s: to.spatial_transformer(torch.from_numpy(onehot), to.rescale_dense_transform(to.vec_int(s, 5), 2)), svf)
    ts = dev(svf.astype(np.float32)).requires_grad_(True)
    model = vxm.networks.VxmDense((8, 12, 16), int_steps=5, svf_resolution=2, int_resolution=2)
    model.deform([np.zeros((1, 8, 12, 16, 1), np.float32), ts])
    pred = vxm.layers.SpatialTransformer(interp_method='linear', name='pred')([dev(onehot.astype(np.float32)), model.references.pos_flow])
    pred.backward(dev(g, 'cl'))
    np.testing.assert_allclose(host(ts.grad), gs, rtol=5e-4, atol=5e-5)


# --------------------------------------------------------------------------------------
# full-size, size-independent properties (BASELINE shapes)
# --------------------------------------------------------------------------------------
def test_full_size_properties():
    torch.manual_seed(0)
    B, X, Y, Z = 2, 80, 80, 96
    # constant SVF integrates to itself; x2 rescale of a constant doubles it; identity warp
    c = torch.tensor([1.25, -0.5, 2.0], device='cuda').expand(B, X, Y, Z, 3).contiguous()
    flow = ops.vecint(c, 7)
    assert torch.equal(ops.to_layout(flow, 'cl'), c)
    up = ops.rescale_dense_transform(flow, 2)
    assert tuple(up.shape) == (B, 160, 160, 192, 3)
    assert torch.allclose(ops.to_layout(up, 'cl'), 2 * c[:, :1, :1, :1].expand(B, 160, 160, 192, 3), rtol=1e-6, atol=0)
    img = torch.rand(B, 160, 160, 192, 1, device='cuda')
    zero = torch.zeros(B, 160, 160, 192, 3, device='cuda')
    assert torch.equal(ops.warp(img, zero), img)
    assert torch.equal(ops.warp(img, zero, 'nearest'), img)
    # layouts agree bit for bit on a random smooth field at full size
    svf = torch.nn.functional.interpolate(torch.randn(B, 3, 10, 10, 12, device='cuda') * 3, size=(X, Y, Z),
                                          mode='trilinear').permute(0, 2, 3, 4, 1).contiguous()
    a = ops.to_layout(ops.vecint(svf, 7), 'cl')
    b = ops.vecint(ops.to_layout(svf, 'planar'), 7, out_layout='cl')
    assert torch.equal(a, b)
    # Jacobian of the zero field is 1, of an affine field det(I + A)
    det, stats = ops.jacobian_determinant(torch.zeros(1, 64, 64, 64, 3, device='cuda'))
    assert torch.all(det == 1) and stats[0, 0].item() == 0 and stats[0, 3].item() == 60 ** 3
