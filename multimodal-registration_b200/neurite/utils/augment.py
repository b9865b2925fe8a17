"""ne.utils.augment mirror: draw_perlin (gen_apply_def_field.py:59, train_synthmorph.py:57,61).

Multi-scale smooth noise: for every scale, Gaussian noise on a coarse grid is up-sampled with
the CUDA resize kernel and summed.  The random stream cannot match TensorFlow's, so parity
with the reference is distributional only; only the resize is part
of the hot path.
"""
import numpy as np
import torch

from ... import _host, ops


def draw_perlin(out_shape, scales, min_std=0, max_std=1, modulate=True, dtype=torch.float32, seeds=None):
    """out_shape = (*spatial, features) with 3 spatial axes, or (X, Y, Z, L, features) as the
    reference passes (the 4th axis is then sampled at full resolution, not across-interpolated)."""
    out_shape = tuple(int(s) for s in out_shape)
    if np.isscalar(scales):
        scales = [scales]
    if len(out_shape) == 5:
        X, Y, Z, L, F = out_shape
        feats = L * F
    elif len(out_shape) == 4:
        X, Y, Z, F = out_shape
        L, feats = None, F
    else:
        raise NotImplementedError('draw_perlin: out_shape must have 3 spatial axes')
    dev = _host.device()
    seeds = dict(seeds) if seeds else {}
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seeds.get('noise', torch.seed() % (2 ** 31))))
    out = None
    for scale in scales:
        sample = tuple(int(np.ceil(d / scale)) for d in (X, Y, Z))
        std = max_std
        if modulate:
            std = min_std + (max_std - min_std) * torch.rand((), generator=gen, device=dev).item()
        gauss = torch.randn((1,) + sample + (feats,), generator=gen, device=dev) * std
        if scale == 1:
            up = gauss
        else:
            zoom = [o / s for o, s in zip((X, Y, Z), sample)]
            up = ops.resize(gauss, zoom)
        out = up if out is None else out + up
    out = ops.to_layout(out, 'cl')[0]
    if L is not None:
        out = out.reshape(X, Y, Z, L, F)
    return out.to(dtype)
