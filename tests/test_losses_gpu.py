"""Dice and Grad('l2') against the torch-CPU restatement, values and gradients."""
import numpy as np
import pytest
import torch

import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
from oracle import torch_oracle as to

pytestmark = pytest.mark.gpu
vxm = mrb.voxelmorph


@pytest.mark.parametrize('C', [1, 2, 5, 26, 40])
@pytest.mark.parametrize('layout', ['cl', 'planar'])
def test_dice_loss_and_gradient(C, layout):
    rng = np.random.default_rng(300 + C)
    shape = (2, 7, 9, 11, C)
    t = rng.random(shape)
    p = rng.random(shape)
    if C > 2:
        t[0, ..., 1] = 0.0                                   # an empty channel in both maps: divide_no_nan -> 0
        p[0, ..., 1] = 0.0
    tt = torch.tensor(t, dtype=torch.float64)
    pp = torch.tensor(p, dtype=torch.float64, requires_grad=True)
    want = to.dice_loss(tt, pp)
    want.backward()
    dt = ops.to_layout(torch.tensor(t, dtype=torch.float32).cuda(), layout)
    dp = ops.to_layout(torch.tensor(p, dtype=torch.float32).cuda(), layout).detach().requires_grad_(True)
    got = ops.dice_loss(dt, dp)
    got.backward()
    np.testing.assert_allclose(float(got), float(want), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(dp.grad.cpu().numpy(), pp.grad.numpy(), rtol=1e-4, atol=1e-9)
    # the mirror of the reference call (train_synthmorph.py:305), numpy in
    np.testing.assert_allclose(float(vxm.losses.Dice().loss(t.astype(np.float32), p.astype(np.float32))), float(want), rtol=1e-5)


@pytest.mark.parametrize('shape', [(6, 8, 10), (5, 7, 9), (2, 2, 2)])
@pytest.mark.parametrize('layout', ['cl', 'planar'])
def test_grad_l2_loss_and_gradient(shape, layout):
    rng = np.random.default_rng(17)
    f = rng.standard_normal((3,) + shape + (3,))
    ff = torch.tensor(f, dtype=torch.float64, requires_grad=True)
    want = to.grad_l2_loss(ff, 0.5)
    w = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    (want * w).sum().backward()
    df = ops.to_layout(torch.tensor(f, dtype=torch.float32).cuda(), layout).detach().requires_grad_(True)
    got = ops.grad_l2_loss(df, 0.5)
    (got * w.float().cuda()).sum().backward()
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(df.grad.cpu().numpy(), ff.grad.numpy(), rtol=1e-4, atol=1e-7)
    got2 = vxm.losses.Grad('l2', loss_mult=0.5).loss(None, f.astype(np.float32))
    np.testing.assert_allclose(got2.cpu().numpy(), want.detach().numpy(), rtol=1e-5)


def test_losses_full_size_channels_last_runs_at_stream_speed():
    # smoke at config.json size (C = 26, 160 x 160 x 192): finite, in range, gradient shape / layout
    torch.manual_seed(0)
    lab = torch.randint(0, 26, (1, 160, 160, 192), device='cuda')
    t = torch.nn.functional.one_hot(lab, 26).float()
    p = torch.rand_like(t).requires_grad_(True)
    loss = ops.dice_loss(t, p)
    loss.backward()
    assert -1.0 <= float(loss) <= 0.0 and p.grad.shape == p.shape and ops.layout_of(p.grad) == 'cl'


@pytest.mark.parametrize('C', [2, 3, 6, 26, 40])
@pytest.mark.parametrize('field_layout,fill', [('cl', None), ('planar', 0.0)])
def test_fused_warp_dice_matches_the_two_ops(C, field_layout, fill):
    rng = np.random.default_rng(700 + C)
    shape = (5, 7, 9)
    img = torch.tensor(rng.random((2,) + shape + (C,)), dtype=torch.float32).cuda()
    true = torch.tensor(rng.random((2,) + shape + (C,)), dtype=torch.float32).cuda()
    f = torch.tensor(rng.standard_normal((2,) + shape + (3,)) * 1.5, dtype=torch.float32).cuda()
    fa = ops.to_layout(f, field_layout).detach().requires_grad_(True)
    fb = ops.to_layout(f, field_layout).detach().requires_grad_(True)
    want = ops.dice_loss(true, ops.warp(img, fa, 'linear', fill))
    want.backward()
    got = ops.warp_dice_loss(img, fb, true, fill)
    got.backward()
    np.testing.assert_allclose(float(got.detach()), float(want.detach()), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(fb.grad.cpu().numpy(), fa.grad.cpu().numpy(), rtol=2e-4, atol=1e-8)


def test_dice_loss_zeropad_as_documented():
    rng = np.random.default_rng(9)
    C, shape = 6, (6, 7, 8)
    lab_t = rng.integers(0, C, (2,) + shape)
    lab_p = rng.integers(0, C, (2,) + shape)
    t = np.eye(C, dtype=np.float64)[lab_t]
    p = np.eye(C, dtype=np.float64)[lab_p] * 0.8 + 0.03          # soft prediction; channel 0 >= 1 never ...
    p[..., 0] = np.where(lab_p == 0, 1.0, p[..., 0])           # ... except where the label is 0 (zero padding)
    tt = torch.tensor(t)
    pp = torch.tensor(p, requires_grad=True)
    want = to.dice_loss_zeropad(tt, pp)
    want.backward()
    dp = torch.tensor(p, dtype=torch.float32).cuda().requires_grad_(True)
    got = vxm.losses.dice_loss_zeropad(torch.tensor(t, dtype=torch.float32).cuda(), dp)
    got.backward()
    np.testing.assert_allclose(float(got.detach()), float(want.detach()), rtol=1e-5)
    np.testing.assert_allclose(dp.grad.cpu().numpy(), pp.grad.numpy(), rtol=1e-4, atol=1e-9)
