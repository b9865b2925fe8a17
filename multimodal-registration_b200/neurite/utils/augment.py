"""ne.utils.augment mirror: draw_perlin (gen_apply_def_field.py:59, train_synthmorph.py:57,61).

Multi-scale smooth noise: for every scale, Gaussian noise on a coarse grid is up-sampled with
the CUDA resize kernel and summed.  The random stream cannot match TensorFlow's, so parity
with the reference is distributional only; only the resize is part
of the hot path.
"""
import numpy as np
import torch

from ... import _coords, _host, ops


def _lerp_axis(vol, axis, n_out):
    """Corner-aligned linear resample of ``vol`` along one axis onto tf.linspace(0, n_in - 1, n_out) (the same
    per-axis set-up as ne.utils.interpn: clip, lower corner i1 - 1, w_lo = i1 - loc).  Multilinear
    interpolation is separable, so a 3-D resize followed by this is the reference's 4-D resize."""
    n_in = vol.shape[axis]
    c = torch.from_numpy(_coords.linspace_tf(n_in, n_out)).to(vol.device)
    cl = c.clamp(0, n_in - 1)
    i1 = (cl.floor().long() + 1).clamp(max=n_in - 1)
    i0 = (i1 - 1).clamp(min=0)
    w0 = i1.to(vol.dtype) - cl
    shape = [1] * vol.dim()
    shape[axis] = n_out
    return vol.index_select(axis, i0) * w0.reshape(shape) + vol.index_select(axis, i1) * (1 - w0).reshape(shape)


def draw_perlin(out_shape, scales, min_std=0, max_std=1, modulate=True, dtype=torch.float32, seeds=None):
    """out_shape = (*spatial, features): 3 spatial axes, or 4 as the reference passes them --
    (X, Y, Z, 1, 3) at gen_apply_def_field.py:59 and (X, Y, Z, 26, 3) at train_synthmorph.py:62, where the
    label axis is a 4th SPATIAL axis: for every scale it is sampled at ceil(L / scale) points and interpolated
    (SURVEY.md Appendix A.11), so at scales >= L every label sees (nearly) the same smooth field."""
    out_shape = tuple(int(s) for s in out_shape)
    if np.isscalar(scales):
        scales = [scales]
    if len(out_shape) == 5:
        X, Y, Z, L, F = out_shape
    elif len(out_shape) == 4:
        X, Y, Z, F = out_shape
        L = None
    else:
        raise NotImplementedError('draw_perlin: out_shape must have 3 or 4 spatial axes')
    dev = _host.device()
    seeds = dict(seeds) if seeds else {}
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seeds.get('noise', torch.seed() % (2 ** 31))))
    out = None
    for scale in scales:
        sample = tuple(int(np.ceil(d / scale)) for d in (X, Y, Z))
        sl = 1 if L is None else int(np.ceil(L / scale))
        std = max_std
        if modulate:
            std = min_std + (max_std - min_std) * torch.rand((), generator=gen, device=dev).item()
        gauss = torch.randn((1,) + sample + (sl * F,), generator=gen, device=dev) * std
        if scale == 1:
            up = gauss
        else:
            zoom = [o / s for o, s in zip((X, Y, Z), sample)]
            up = ops.to_layout(ops.resize(gauss, zoom), 'cl')                       # the hot-path resize kernel, per coarse label slice
        if L is not None:
            up = up.reshape(up.shape[1:4] + (sl, F))
            if scale != 1:
                up = _lerp_axis(up, 3, int(sl * (L / sl)))                          # the label axis, like the other three
        else:
            up = up[0]
        out = up if out is None else out + up
    return out.to(dtype)
