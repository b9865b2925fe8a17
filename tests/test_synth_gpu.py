"""ne.models.labels_to_image mirror (train_synthmorph.py:258-289): the random stream cannot match TensorFlow's, so the
generator is checked structurally and distributionally, and each kernel of the intensity model against NumPy."""
import numpy as np
import pytest
import torch

import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import _lib
from multimodal_registration_b200.neurite import models
from multimodal_registration_b200.ops import _ptr, _stream

pytestmark = pytest.mark.gpu


def test_synth_intensity_per_label_statistics_and_determinism():
    n, L = 400_003, 5
    rng = np.random.default_rng(0)
    lab = torch.from_numpy(rng.integers(0, L, n).astype(np.float32)).cuda()
    means = torch.tensor([10., 50., 90., 130., 200.], device='cuda')
    stds = torch.tensor([1., 5., 10., 2., 20.], device='cuda')
    out = torch.empty(n, device='cuda')
    _lib.call('dfm_synth_intensity', _ptr(lab), _ptr(means), _ptr(stds), L, 1234, _ptr(out), n, _stream())
    out2 = torch.empty(n, device='cuda')
    _lib.call('dfm_synth_intensity', _ptr(lab), _ptr(means), _ptr(stds), L, 1234, _ptr(out2), n, _stream())
    assert torch.equal(out, out2)                                    # a pure function of (seed, voxel)
    _lib.call('dfm_synth_intensity', _ptr(lab), _ptr(means), _ptr(stds), L, 1235, _ptr(out2), n, _stream())
    assert not torch.equal(out, out2)
    o, l = out.cpu().numpy(), lab.cpu().numpy().astype(int)
    z_all = []
    for k in range(L):
        v = o[l == k]
        m = v.size
        assert abs(v.mean() - means[k].item()) < 5 * stds[k].item() / np.sqrt(m)
        assert abs(v.std() / stds[k].item() - 1) < 0.02
        z_all.append((v - means[k].item()) / stds[k].item())
    z = np.concatenate(z_all)
    # normality: skewness ~ 0, kurtosis ~ 3, tails present, no correlation between neighbouring voxels
    assert abs((z ** 3).mean()) < 0.02 and abs((z ** 4).mean() - 3) < 0.05 and np.abs(z).max() > 4
    zz = (o - means.cpu().numpy()[l]) / stds.cpu().numpy()[l]
    assert abs(np.corrcoef(zz[:-1], zz[1:])[0, 1]) < 0.01 and abs(np.corrcoef(zz[:-4], zz[4:])[0, 1]) < 0.01


@pytest.mark.parametrize('axis', [0, 1, 2])
def test_conv1d_axis_matches_numpy(axis):
    rng = np.random.default_rng(axis)
    a = rng.random((2, 7, 9, 11)).astype(np.float32)
    taps = models.gaussian_taps(0.8, 1.0)
    out = torch.empty(a.shape, device='cuda')
    ta, tt = torch.from_numpy(a).cuda(), torch.from_numpy(taps).cuda()           # named: they must outlive the call
    _lib.call('dfm_conv1d_axis', _ptr(ta), _ptr(out), 2, 7, 9, 11, axis, _ptr(tt), taps.size, _stream())
    want = np.apply_along_axis(lambda r: np.convolve(r, taps[::-1], mode='same'), axis + 1, a.astype(np.float64))
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-5, atol=1e-6)
    assert abs(taps.sum() - 1) < 1e-6 and taps.size == 7


def test_norm_gamma_exp_clip_and_onehot():
    rng = np.random.default_rng(3)
    img = (rng.random((2, 1000)) * 300 - 20).astype(np.float32)
    bias = (rng.standard_normal((2, 1000)) * 0.3).astype(np.float32)
    t = torch.from_numpy(img).cuda()
    out = torch.empty_like(t)
    tb = torch.from_numpy(bias).cuda()
    _lib.call('dfm_scale_exp_clip', _ptr(t), _ptr(tb), _ptr(out), t.numel(), 0.0, 255.0, _stream())
    want = np.clip(img * np.exp(bias), 0, 255)
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=2e-6, atol=1e-4)
    mm = torch.tensor([[want[0].min(), want[0].max()], [want[1].min(), want[1].max()]], device='cuda', dtype=torch.float64)
    gamma = torch.tensor([0.7, 1.4], device='cuda')
    res = torch.empty_like(out)
    _lib.call('dfm_norm_gamma', _ptr(out), _ptr(mm), _ptr(gamma), _ptr(res), 2, 1000, _stream())
    w = np.stack([((want[b] - want[b].min()) / (want[b].max() - want[b].min())) ** g for b, g in enumerate((0.7, 1.4))])
    np.testing.assert_allclose(res.cpu().numpy(), w, rtol=2e-5, atol=2e-6)
    lab = torch.tensor([0., 3., 5., 2., 9., -1.], device='cuda')
    lut = torch.tensor([0, -1, 1, 2, -1, 3], device='cuda', dtype=torch.int32)          # labels 1, 4 dropped; 9 and -1 outside
    oh = torch.empty((6, 4), device='cuda')
    _lib.call('dfm_onehot', _ptr(lab), _ptr(lut), 6, 4, _ptr(oh), 6, _stream())
    want_oh = np.zeros((6, 4), np.float32)
    want_oh[0, 0] = want_oh[1, 2] = want_oh[2, 3] = want_oh[3, 1] = 1
    np.testing.assert_array_equal(oh.cpu().numpy(), want_oh)


def test_labels_to_image_generator():
    """The call of train_synthmorph.py:258-289 with config.json's generator arguments at a reduced shape."""
    shape, L = (32, 32, 48), 6
    rng = np.random.default_rng(1)
    coarse = rng.integers(0, L, (4, 4, 6))
    labels = np.kron(coarse, np.ones((8, 8, 8), dtype=np.int64)).astype(np.float32)[None, ..., None]     # blocky label map
    gen_args = dict(in_shape=shape, in_label_list=np.arange(L), out_label_list=np.arange(L), warp_std=3, warp_res=16,
                    blur_std=1, bias_std=0.3, bias_res=40, gamma_std=0.25)
    gen = mrb.neurite.models.labels_to_image(**gen_args, id=0, seeds={'all': 5})
    image, onehot = gen(np.repeat(labels, 2, 0))
    assert tuple(image.shape) == (2,) + shape + (1,) and tuple(onehot.shape) == (2,) + shape + (L,)
    im, oh = image.cpu().numpy(), onehot.cpu().numpy()
    assert im.min() == 0.0 and im.max() == 1.0 and np.isfinite(im).all()             # min-max normalised per item
    assert set(np.unique(oh)) <= {0.0, 1.0} and (oh.sum(-1) <= 1).all()
    assert (oh.sum(-1) == 0).mean() < 0.5                                            # fill_value 0 voxels map to label 0, not to nothing
    # the two items are deformed differently, and differently from the input
    lab0, lab1 = oh[0].argmax(-1), oh[1].argmax(-1)
    assert (lab0 != lab1).mean() > 0.01 and (lab0 != labels[0, ..., 0]).mean() > 0.01
    # intensities follow the labels: the between-label variance of the image dominates the within-label variance
    means = np.array([im[0, ..., 0][lab0 == k].mean() for k in range(L) if (lab0 == k).sum() > 100])
    within = np.mean([im[0, ..., 0][lab0 == k].std() for k in range(L) if (lab0 == k).sum() > 100])
    assert means.std() > within * 0.5
    # same seed -> same sample; another id -> another sample
    image_b, _ = mrb.neurite.models.labels_to_image(**gen_args, id=0, seeds={'all': 5})(np.repeat(labels, 2, 0))
    assert torch.equal(image, image_b)
    image_c, _ = mrb.neurite.models.labels_to_image(**gen_args, id=1, seeds={'all': 6})(np.repeat(labels, 2, 0))
    assert not torch.equal(image, image_c)
    # out_label_list as a subset: dropped labels vanish from the one-hot map
    sub = mrb.neurite.models.labels_to_image(**dict(gen_args, out_label_list=[1, 3]), seeds={'all': 5})
    _, oh2 = sub.predict(labels)
    assert oh2.shape[-1] == 2 and oh2.sum() > 0 and (oh2.sum(-1) == 0).any()


def test_generator_onehot_map_is_warped_from_its_label_map():
    """pred = SpatialTransformer('linear')([map_1, flow]) (train_synthmorph.py:298) on the generator's one-hot output takes the
    label-map kernel (ops.warp_onehot); a plain copy of the same tensor takes the generic channels-last kernel: same bits,
    same field gradient, also with an out_label_list that drops labels (all-zero rows)."""
    rng = np.random.default_rng(9)
    shape = (16, 24, 32)
    labels = rng.integers(0, 6, (2,) + shape + (1,)).astype(np.float32)
    for out_list in (None, [1, 3, 4]):
        gen = mrb.neurite.models.labels_to_image(in_shape=shape, in_label_list=list(range(6)), out_label_list=out_list, warp_std=2.0,
                                                 warp_res=[8], blur_std=1.0, bias_std=0.3, bias_res=[16], gamma_std=0.1, seeds={'all': 3})
        _, onehot = gen(labels)
        assert hasattr(onehot, 'dfm_labels') and onehot.dfm_labels[1] == onehot.shape[-1]
        flow = (rng.standard_normal((2,) + shape + (3,)) * 1.5).astype(np.float32)
        layer = mrb.voxelmorph.layers.SpatialTransformer(interp_method='linear')
        f1 = torch.from_numpy(flow).cuda().requires_grad_(True)
        f2 = torch.from_numpy(flow).cuda().requires_grad_(True)
        fast = layer([onehot, f1])
        plain = layer([onehot.clone(), f2])                       # a copy does not carry the label map
        assert torch.equal(fast, plain)
        g = torch.from_numpy(rng.standard_normal(tuple(fast.shape)).astype(np.float32)).cuda()
        fast.backward(g)
        plain.backward(g)
        assert torch.allclose(f1.grad, f2.grad, rtol=1e-4, atol=1e-5 * float(f2.grad.abs().max()))
