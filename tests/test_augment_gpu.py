"""ne.utils.augment.draw_perlin (gen_apply_def_field.py:59, train_synthmorph.py:57-64): the random stream cannot
match TensorFlow's, so these are structural / distributional checks plus an oracle check of the 4-D resize it is
built from (the label axis of a 5-entry shape is a 4th spatial axis)."""
import numpy as np
import pytest
import torch

import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
from multimodal_registration_b200.neurite.utils import augment
from oracle import interp_oracle as io

pytestmark = pytest.mark.gpu


def test_four_d_resize_matches_oracle():
    """3-D resize kernel per coarse label slice + a lerp along the label axis == the reference's 4-D ne.utils.resize."""
    rng = np.random.default_rng(0)
    vol = rng.standard_normal((3, 4, 5, 2, 3)).astype(np.float32)             # (sx, sy, sz, sL, F)
    zoom = [4.0, 3.0, 2.4, 13.0]
    want = io.resize(vol, zoom)                                               # 16 corners, 4-D
    g = torch.from_numpy(vol).cuda().reshape(1, 3, 4, 5, 6)
    up = ops.to_layout(ops.resize(g, zoom[:3]), 'cl')
    up = up.reshape(up.shape[1:4] + (2, 3))
    got = augment._lerp_axis(up, 3, int(2 * zoom[3])).cpu().numpy()
    assert got.shape == want.shape == (12, 12, 12, 26, 3)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


def test_label_axis_is_sampled_coarsely_and_interpolated():
    # train_synthmorph.py:61-64: 26 labels, scales >= 26 -> one sample along the label axis: every label gets the same field
    f = augment.draw_perlin((32, 32, 48, 26, 3), scales=(32, 64), max_std=3, seeds={'noise': 1})
    assert tuple(f.shape) == (32, 32, 48, 26, 3)
    assert torch.equal(f[..., 0, :], f[..., 25, :])
    # scale 16: two samples along the label axis, linearly interpolated -> neighbouring labels nearly identical,
    # and every label is the exact blend of the two end labels
    f = augment.draw_perlin((32, 32, 48, 26, 3), scales=16, max_std=3, modulate=False, seeds={'noise': 2})
    a, b, m = f[..., 0, :].double(), f[..., 25, :].double(), f[..., 10, :].double()
    corr = ((a - a.mean()) * (f[..., 1, :].double() - f[..., 1, :].double().mean())).mean() / (a.std() * f[..., 1, :].double().std())
    assert corr > 0.99
    torch.testing.assert_close(m, a + (b - a) * (10.0 / 25.0), rtol=1e-4, atol=1e-4)
    assert not torch.equal(a, b)


def test_singleton_label_axis_and_std():
    # gen_apply_def_field.py:59: out_shape (*shape, 1, 3); scale 1 is plain Gaussian noise of std max_std
    f = augment.draw_perlin((24, 24, 32, 1, 3), scales=(1,), max_std=2.0, modulate=False, seeds={'noise': 3})
    assert tuple(f.shape) == (24, 24, 32, 1, 3)
    assert abs(f.std().item() - 2.0) < 0.05
    g = augment.draw_perlin((24, 24, 32, 3), scales=(8, 16), max_std=1.0, seeds={'noise': 4})
    assert tuple(g.shape) == (24, 24, 32, 3)
    # smooth: neighbouring voxels are strongly correlated at scale >= 8
    d = (g[1:] - g[:-1]).std().item()
    assert d < 0.3 * g.std().item()
    # the same seed reproduces the field
    assert torch.equal(g, augment.draw_perlin((24, 24, 32, 3), scales=(8, 16), max_std=1.0, seeds={'noise': 4}))
