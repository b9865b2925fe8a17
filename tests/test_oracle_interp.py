"""Anchor the (unpinned) interpolation oracle with oracle-independent known answers and an
independent float64 implementation (scipy map_coordinates, order=1, mode='nearest')."""
import numpy as np
import pytest
import scipy.ndimage as ndi
import torch

from oracle import interp_oracle as io
from oracle import torch_oracle as to

RNG = np.random.default_rng(7)


def rand_field(shape, std):
    return (RNG.standard_normal(shape + (3,)) * std).astype(np.float32)


def test_linspace_tf():
    l = io.linspace_tf(0., 79., 160)
    assert l.dtype == np.float32 and l.shape == (160,)
    assert l[0] == 0 and l[-1] == 79
    delta = np.float32(np.float32(79.) / np.float32(159.))
    assert l[7] == np.float32(delta * np.float32(7))
    # identity grid: delta == 1 exactly
    np.testing.assert_array_equal(io.linspace_tf(0., 9., 10), np.arange(10, dtype=np.float32))
    np.testing.assert_array_equal(io.linspace_tf(0., 4., 1), np.array([0.], np.float32))


def test_identity_warp_is_exact():
    vol = RNG.random((6, 7, 8, 2)).astype(np.float32)
    z = np.zeros((6, 7, 8, 3), np.float32)
    np.testing.assert_array_equal(io.transform(vol, z), vol)
    np.testing.assert_array_equal(io.transform(vol, z, 'nearest'), vol)


def test_integer_translation_with_edge_clamp():
    vol = RNG.random((6, 7, 8, 1)).astype(np.float32)
    s = np.zeros((6, 7, 8, 3), np.float32)
    s[..., 0] = 2
    s[..., 2] = -3
    out = io.transform(vol, s)
    xi = np.clip(np.arange(6) + 2, 0, 5)
    zi = np.clip(np.arange(8) - 3, 0, 7)
    np.testing.assert_array_equal(out, vol[xi][:, :, zi])


def test_linear_ramp_known_answer():
    X, Y, Z = 9, 10, 11
    g = np.stack(np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing='ij'), -1)
    coef = np.array([0.5, -1.25, 2.0])
    vol = (g @ coef + 3.0).astype(np.float32)[..., None]
    shift = (RNG.random((X, Y, Z, 3)) * 1.5 - 0.75).astype(np.float32)
    loc = g + shift.astype(np.float64)
    inb = np.all((loc >= 0) & (loc <= np.array([X - 1, Y - 1, Z - 1])), -1)
    out = io.transform(vol, shift)[..., 0]
    expect = loc @ coef + 3.0
    np.testing.assert_allclose(out[inb], expect[inb], rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize('std', [0.3, 3.0, 12.0])
def test_against_scipy_map_coordinates(std):
    X, Y, Z = 10, 12, 9
    vol = RNG.random((X, Y, Z, 2)).astype(np.float32)
    shift = rand_field((X, Y, Z), std)
    out = io.transform(vol, shift)
    g = np.stack(np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing='ij'), 0).astype(np.float64)
    loc = g + np.moveaxis(shift.astype(np.float64), -1, 0)
    # edge clamp == clip the sampling location into the volume, then plain trilinear
    for d, n in enumerate((X, Y, Z)):
        loc[d] = np.clip(loc[d], 0, n - 1)
    for c in range(2):
        ref = ndi.map_coordinates(vol[..., c].astype(np.float64), loc, order=1, mode='nearest')
        np.testing.assert_allclose(out[..., c], ref, rtol=1e-5, atol=2e-6)


def test_nearest_round_half_even_and_clamp():
    vol = np.arange(8, dtype=np.float32).reshape(8, 1, 1, 1)
    shift = np.zeros((8, 1, 1, 3), np.float32)
    shift[:, 0, 0, 0] = [0.5, 0.5, 0.5, -0.5, 1.5, -9, 9, 0.49]
    out = io.transform(vol, shift, 'nearest')[:, 0, 0, 0]
    # loc = [0.5, 1.5, 2.5, 2.5, 5.5, -4, 15, 7.49] -> rint -> [0, 2, 2, 2, 6, -4, 15, 7] -> clamp
    np.testing.assert_array_equal(out, [0, 2, 2, 2, 6, 0, 7, 7])


def test_fill_value_uses_unclipped_location():
    vol = np.ones((4, 4, 4, 1), np.float32)
    shift = np.zeros((4, 4, 4, 3), np.float32)
    shift[0, :, :, 0] = -0.25        # loc < 0  -> fill
    shift[3, :, :, 0] = 0.25         # loc > max -> fill
    shift[1, :, :, 0] = 0.5          # in bounds, fractional
    out = io.transform(vol, shift, fill_value=-7.0)[..., 0]
    assert np.all(out[0] == -7) and np.all(out[3] == -7) and np.all(out[1] == 1) and np.all(out[2] == 1)
    outn = io.transform(vol, shift, 'nearest', fill_value=0)[..., 0]
    assert np.all(outn[0] == 0) and np.all(outn[3] == 0) and np.all(outn[1] == 1)


def test_constant_svf_integrates_to_itself():
    c = np.array([1.5, -0.75, 0.25], np.float32)
    svf = np.broadcast_to(c, (6, 6, 6, 3)).copy()
    out = io.integrate_vec(svf, 7)
    np.testing.assert_array_equal(out, svf)      # power-of-two scaling: exact
    np.testing.assert_array_equal(io.integrate_vec(svf, 0), svf)


def test_rescale_constant_and_identity_and_shapes():
    c = np.broadcast_to(np.array([1., -2., 3.], np.float32), (4, 5, 6, 3)).copy()
    up = io.rescale_dense_transform(c, 2)
    assert up.shape == (8, 10, 12, 3)
    np.testing.assert_allclose(up, 2 * np.broadcast_to(c[0, 0, 0], up.shape), rtol=1e-6)
    f = rand_field((4, 5, 6), 2.0)
    np.testing.assert_array_equal(io.rescale_dense_transform(f, 1), f)      # identity grid
    dn = io.rescale_dense_transform(rand_field((8, 10, 12), 1.0), 0.5)
    assert dn.shape == (4, 5, 6, 3)
    b = io.rescale_dense_transform(np.stack([f, f]), 2)
    assert b.shape == (2, 8, 10, 12, 3)
    np.testing.assert_array_equal(b[0], io.rescale_dense_transform(f, 2))
    # corner alignment: the first/last samples of the upsampled grid are the input corners * 2
    u = io.rescale_dense_transform(f, 2)
    np.testing.assert_array_equal(u[0, 0, 0], 2 * f[0, 0, 0])
    np.testing.assert_array_equal(u[-1, -1, -1], 2 * f[-1, -1, -1])


def test_compose_identities():
    a = rand_field((5, 6, 7), 1.0)
    z = np.zeros_like(a)
    np.testing.assert_array_equal(io.compose([a, z]), a)      # then zero
    np.testing.assert_array_equal(io.compose([z, a]), a)      # zero first
    with pytest.raises(ValueError):
        io.compose([a])
    # constant translations add
    t1 = np.broadcast_to(np.array([1., 0., -2.], np.float32), a.shape)
    t2 = np.broadcast_to(np.array([0.5, 0.25, 1.], np.float32), a.shape)
    np.testing.assert_array_equal(io.compose([t1, t2]), t1 + t2)


def test_channelwise_equals_per_channel_shared():
    X, Y, Z, C = 5, 6, 7, 3
    vol = RNG.random((X, Y, Z, C)).astype(np.float32)
    shift = (RNG.standard_normal((X, Y, Z, C, 3)) * 2).astype(np.float32)
    out = io.transform(vol, shift)
    assert out.shape == (X, Y, Z, C)
    for c in range(C):
        np.testing.assert_array_equal(out[..., c], io.transform(vol[..., c:c + 1], shift[..., c, :])[..., 0])


def test_transform_output_takes_field_shape():
    vol = RNG.random((6, 6, 6, 1)).astype(np.float32)
    shift = rand_field((4, 5, 3), 1.0)
    assert io.transform(vol, shift).shape == (4, 5, 3, 1)
    with pytest.raises(ValueError):
        io.transform(vol, np.zeros((6, 6, 6, 2), np.float32))


def test_torch_oracle_bitwise_equals_numpy_oracle():
    X, Y, Z = 8, 9, 10
    vol = RNG.random((X, Y, Z, 2)).astype(np.float32)
    shift = rand_field((X, Y, Z), 3.0)
    for method in ('linear', 'nearest'):
        for fv in (None, 0.5):
            a = io.transform(vol, shift, method, fill_value=fv)
            b = to.transform(torch.from_numpy(vol), torch.from_numpy(shift), method, fill_value=fv).numpy()
            np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(io.integrate_vec(shift, 5), to.integrate_vec(torch.from_numpy(shift), 5).numpy())
    np.testing.assert_array_equal(io.rescale_dense_transform(shift, 2),
                                  to.rescale_dense_transform(torch.from_numpy(shift), 2).numpy())
    np.testing.assert_array_equal(io.rescale_dense_transform(shift, 0.5),
                                  to.rescale_dense_transform(torch.from_numpy(shift), 0.5).numpy())
    s2 = rand_field((X, Y, Z), 1.0)
    np.testing.assert_array_equal(io.compose([shift, s2]),
                                  to.compose([torch.from_numpy(shift), torch.from_numpy(s2)]).numpy())


def test_torch_oracle_gradcheck_fp64():
    torch.manual_seed(0)
    vol = torch.rand(4, 5, 3, 2, dtype=torch.float64, requires_grad=True)
    shift = (torch.randn(4, 5, 3, 3, dtype=torch.float64) * 0.7).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda v, s: to.transform(v, s), (vol, shift), eps=1e-6, atol=1e-5)
