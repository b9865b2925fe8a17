"""ne.utils mirror: interpn, resize / zoom (+ augment.draw_perlin)."""
import torch

from ... import _host, ops
from . import augment   # noqa: F401


def interpn(vol, loc, interp_method='linear', fill_value=None):
    """N-D (here 3-D) gather interpolation of ``vol [X, Y, Z(, C)]`` at absolute locations
    ``loc [..., 3]`` (or a list of 3 arrays); edge clamp, optional fill_value."""
    vol = _host.to_device(vol)
    if isinstance(loc, (list, tuple)):
        loc = torch.stack([_host.to_device(l, torch.float32) for l in loc], -1)
    else:
        loc = _host.to_device(loc, torch.float32)
    nb_dims = loc.shape[-1]
    if nb_dims != 3:
        raise NotImplementedError('interpn: only 3-D volumes are supported, got %d-D' % nb_dims)
    if vol.dim() not in (nb_dims, nb_dims + 1):
        raise ValueError('Number of loc Tensors %d does not match volume dimension %d' % (nb_dims, vol.dim() - 1))
    if vol.dim() == nb_dims:
        vol = vol[..., None]
    grid = tuple(loc.shape[:-1])
    g3 = grid[-3:] if len(grid) >= 3 else (1,) * (3 - len(grid)) + grid
    lead = 1
    for s in grid[:-3]:
        lead *= int(s)
    g3 = (g3[0] * lead,) + tuple(g3[1:])
    out = ops.warp(vol[None], loc.reshape((1,) + g3 + (3,)), interp_method, fill_value, loc_absolute=True)
    return out.reshape(grid + (vol.shape[-1],))


def resize(vol, zoom_factor, interp_method='linear'):
    """Corner-aligned resample of an unbatched ``vol [X, Y, Z, C]`` onto
    ``linspace(0, n-1, int(n*zoom))`` per axis."""
    vol = _host.to_device(vol, torch.float32)
    if vol.dim() != 4:
        raise NotImplementedError('resize: expected an unbatched [X, Y, Z, C] volume')
    if isinstance(zoom_factor, (list, tuple)) and len(zoom_factor) != 3:
        raise NotImplementedError('resize: zoom_factor must have 3 entries')
    return ops.resize(vol[None], zoom_factor, interp_method)[0]


zoom = resize
