#!/usr/bin/env python
"""Randomised parity sweep against the oracle (a tuning / debugging aid, not part of the test suite):
random shapes (odd sizes, Z not a multiple of 4, 2-voxel axes), batch sizes, layouts, channel counts,
fill values and step counts for warp / VecInt / rescale / compose / Jacobian, in both builds.

  [FUZZ_MAXDIM=20] python scripts/fuzz_parity.py [n_cases] [seed]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops
from oracle import interp_oracle as io
from oracle import jacobian_oracle as jo

RTOL, ATOL = 1e-5, 1e-4


def dev(a, layout='cl'):
    return ops.to_layout(torch.as_tensor(a).cuda(), layout)


def host(t):
    return ops.to_layout(t, 'cl').cpu().numpy()


def check(name, got, want, exact):
    if exact:
        ok = np.array_equal(got, want)
    else:
        ok = np.allclose(got, want, rtol=RTOL, atol=ATOL)
    if not ok:
        bad = np.abs(got.astype(np.float64) - want.astype(np.float64))
        print('MISMATCH', name, 'max abs', bad.max(), 'n', int((bad > ATOL).sum()))
    return ok


def smooth(rng, shape, std):
    c = rng.standard_normal(shape).astype(np.float32)
    for ax in (1, 2, 3):
        c = (c + np.roll(c, 1, ax) + np.roll(c, -1, ax)) / 3
    return (c / max(c.std(), 1e-6) * std).astype(np.float32)


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    fails = 0
    for exact in (False, True):
        mrb._lib.use(exact)
        for case in range(n_cases):
            B = int(rng.integers(1, 4))
            maxd = int(os.environ.get('FUZZ_MAXDIM', '20'))
            shape = tuple(int(rng.integers(2, maxd + 1)) for _ in range(3))
            if rng.random() < 0.5:
                shape = shape[:2] + (int(rng.integers(1, maxd // 4 + 2)) * 4,)   # TMA-eligible Z
            std = float(rng.choice([0.05, 0.5, 2.0, 6.0]))
            fl = str(rng.choice(['cl', 'planar']))
            il = str(rng.choice(['cl', 'planar']))
            C = int(rng.choice([1, 1, 2, 3, 7, 26, 33, 40]))
            fill = None if rng.random() < 0.6 else float(rng.choice([0.0, -1.5]))
            tag = 'B%d %s std%.2f C%d %s/%s fill=%s exact=%d' % (B, shape, std, C, il, fl, fill, exact)
            field = smooth(rng, (B,) + shape + (3,), std)
            img = rng.random((B,) + shape + (C,)).astype(np.float32)
            for interp in ('linear', 'nearest'):
                want = io.spatial_transformer(img, field, interp, fill)
                got = host(ops.warp(dev(img, il), dev(field, fl), interp, fill))
                fails += not check('warp %s %s' % (interp, tag), got, want, exact or interp == 'nearest')
            nsteps = int(rng.integers(0, 8))
            want = io.vec_int(field, nsteps)
            got = host(ops.vecint(dev(field, fl), nsteps, out_layout=str(rng.choice(['cl', 'planar']))))
            fails += not check('vecint n=%d %s' % (nsteps, tag), got, want, exact)
            other = smooth(rng, (B,) + shape + (3,), std * 0.5)
            want = np.stack([io.compose([field[i], other[i]]) for i in range(B)])
            fails += not check('compose %s' % tag, host(ops.compose([dev(field, fl), dev(other, il)])), want, exact)
            factor = float(rng.choice([2, 0.5, 1.5, 1, 3]))
            if min(int(s * factor) for s in shape) >= 1:
                want = io.rescale_dense_transform(field, factor)
                fails += not check('rescale x%g %s' % (factor, tag), host(ops.rescale_dense_transform(dev(field, fl), factor)), want, exact)
            if factor >= 1 and min(int(s * factor) for s in shape) >= 1:
                # fused RescaleTransform + warp of a one-channel image (texture-gather / marching kernels where the shapes allow,
                # the brick fusion or the two-kernel fallback otherwise): against the chained oracle and the unfused kernels
                full = tuple(int(s * factor) for s in shape)
                ishape = full if rng.random() < 0.7 else tuple(int(rng.integers(2, maxd + 1)) for _ in range(3))
                img1 = rng.random((B,) + ishape + (1,)).astype(np.float32)
                lab1 = rng.integers(0, 9, (B,) + ishape + (1,)).astype(np.float32)
                d_half = dev(field, 'planar')
                up = ops.rescale_dense_transform(d_half, factor)
                want = io.spatial_transformer(img1, io.rescale_dense_transform(field, factor), 'linear', fill)
                got = host(ops.rescale_warp(dev(img1), d_half, factor, fill))
                if fill is None or exact:
                    fails += not check('fused rescale+warp x%g %s' % (factor, tag), got, want, exact)
                elif np.mean(~np.isclose(got, want, rtol=RTOL, atol=ATOL)) >= 1e-3:      # a few-ulp flow change can flip a voxel across the fill boundary
                    print('MISMATCH fused rescale+warp (fill)', tag)
                    fails += 1
                got_nn = host(ops.rescale_warp(dev(lab1), d_half, factor, fill, 'nearest'))
                fails += not check('fused rescale+nearest x%g %s' % (factor, tag), got_nn, host(ops.warp(dev(lab1), up, 'nearest', fill)), True)
                if exact:
                    fails += not check('fused rescale+nearest vs oracle x%g %s' % (factor, tag), got_nn,
                                       io.spatial_transformer(lab1, io.rescale_dense_transform(field, factor), 'nearest', fill), True)
            if min(shape) >= 5:
                det, stats = ops.jacobian_determinant(dev(field, fl), out_dtype=torch.float64)
                want = np.stack([jo.jacobian_determinant(field[i][:, :, :, None, :].astype(np.float64))[0] for i in range(B)])
                if not np.allclose(det.cpu().numpy().reshape(want.shape), want, rtol=1e-4, atol=1e-4):
                    print('MISMATCH jacdet', tag)
                    fails += 1
    print('fuzz done: %d mismatches' % fails)
    return 1 if fails else 0


if __name__ == '__main__':
    sys.exit(main())
