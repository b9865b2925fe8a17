"""torch-CPU restatement of the same ops as ``interp_oracle`` -- TEST ORACLE ONLY.

PARITY UNPINNED (see ``oracle/__init__.py``).  Same formulas, same op order, written on
torch tensors so that (a) autograd yields the gradients TensorFlow's autodiff would give
for the reference graph (floor: zero gradient; clip: pass-through inside [0, max], inclusive;
gather: scatter-add) -- the oracle for the backward kernels -- and (b) it runs on all host
threads, which makes it the "restated reference" CPU baseline of ``bench.py``
(BASELINE.md section 5).  dtype follows the inputs (fp32 for parity, fp64 for gradcheck).
"""
import itertools

import torch

from . import interp_oracle as _np_oracle


def _linspace(n_in, n_out, dtype):
    # the sampling grid is a constant of the graph: reuse the fp32 tf.linspace restatement
    return torch.from_numpy(_np_oracle.linspace_tf(0., n_in - 1., n_out)).to(dtype)


def interpn(vol, loc, interp_method='linear', fill_value=None):
    if isinstance(loc, (list, tuple)):
        loc = torch.stack(list(loc), -1)
    nb_dims = loc.shape[-1]
    if vol.dim() == nb_dims:
        vol = vol[..., None]
    dt = vol.dtype if vol.dtype.is_floating_point else torch.float32
    loc = loc.to(dt)
    volshape = vol.shape[:-1]
    max_loc = [d - 1 for d in volshape]
    vol_flat = vol.reshape(-1, vol.shape[-1])
    strides = [1] * nb_dims
    for d in range(nb_dims - 2, -1, -1):
        strides[d] = strides[d + 1] * volshape[d + 1]

    if interp_method == 'linear':
        vol_flat = vol_flat.to(dt)
        loc0 = torch.floor(loc)
        clipped = [loc[..., d].clamp(0, max_loc[d]) for d in range(nb_dims)]
        loc0lst = [loc0[..., d].clamp(0, max_loc[d]) for d in range(nb_dims)]
        loc1 = [(loc0lst[d] + 1).clamp(0, max_loc[d]) for d in range(nb_dims)]
        locs = [[f.to(torch.int64) for f in loc0lst], [f.to(torch.int64) for f in loc1]]
        diff_loc1 = [loc1[d].detach() - clipped[d] for d in range(nb_dims)]
        diff_loc0 = [1 - d for d in diff_loc1]
        weights_loc = [diff_loc1, diff_loc0]
        out = 0
        for c in itertools.product([0, 1], repeat=nb_dims):
            idx = 0
            for d in range(nb_dims):
                idx = idx + locs[c[d]][d] * strides[d]
            vol_val = vol_flat[idx.reshape(-1)].reshape(idx.shape + (vol.shape[-1],))
            wt = weights_loc[c[0]][0]
            for d in range(1, nb_dims):
                wt = wt * weights_loc[c[d]][d]
            out = out + wt[..., None] * vol_val
    elif interp_method == 'nearest':
        r = torch.round(loc).to(torch.int64)          # half-to-even, unclipped
        idx = 0
        for d in range(nb_dims):
            idx = idx + r[..., d].clamp(0, max_loc[d]) * strides[d]
        out = vol_flat[idx.reshape(-1)].reshape(idx.shape + (vol.shape[-1],))
    else:
        raise ValueError(interp_method)

    if fill_value is not None:
        oob = torch.zeros(loc.shape[:-1], dtype=torch.bool)
        for d in range(nb_dims):
            oob = oob | (loc[..., d] < 0) | (loc[..., d] > max_loc[d])
        oob = oob[..., None]
        out = out * (~oob).to(out.dtype)
        out = out + oob.to(out.dtype) * torch.as_tensor(fill_value, dtype=out.dtype)
    return out


def resize(vol, zoom_factor, interp_method='linear'):
    if isinstance(zoom_factor, (list, tuple)):
        ndims = len(zoom_factor)
        zoom = list(zoom_factor)
        vol_shape = vol.shape[:ndims]
    else:
        vol_shape = vol.shape[:-1]
        ndims = len(vol_shape)
        zoom = [zoom_factor] * ndims
    new_shape = [int(vol_shape[d] * zoom[d]) for d in range(ndims)]
    dt = vol.dtype if vol.dtype.is_floating_point else torch.float32
    lin = [_linspace(vol_shape[d], new_shape[d], dt) for d in range(ndims)]
    grid = torch.meshgrid(*lin, indexing='ij')
    return interpn(vol, list(grid), interp_method)


def transform(vol, loc_shift, interp_method='linear', indexing='ij', fill_value=None):
    if indexing != 'ij':
        raise ValueError('ij only')
    loc_volshape = loc_shift.shape[:-1]
    nb_dims = vol.dim() - 1
    is_channelwise = len(loc_volshape) == nb_dims + 1
    if loc_shift.shape[-1] != nb_dims:
        raise ValueError('dimension mismatch')
    mesh = torch.meshgrid(*[torch.arange(int(d)) for d in loc_volshape], indexing='ij')
    loc = [mesh[d].to(loc_shift.dtype) + loc_shift[..., d] for d in range(nb_dims)]
    if is_channelwise:
        loc.append(mesh[-1].to(loc_shift.dtype))
    out = interpn(vol, loc, interp_method, fill_value)
    return out[..., 0] if is_channelwise else out


def integrate_vec(vec, nb_steps):
    vec = vec / (2 ** nb_steps)
    for _ in range(nb_steps):
        vec = vec + transform(vec, vec)
    return vec


def rescale_dense_transform(trf, factor, interp_method='linear'):
    if trf.dim() > trf.shape[-1] + 1:
        return torch.stack([rescale_dense_transform(t, factor, interp_method) for t in trf], 0)
    if factor < 1:
        return resize(trf, factor, interp_method) * factor
    return resize(trf * factor, factor, interp_method)


def compose(transforms, interp_method='linear'):
    if len(transforms) < 2:
        raise ValueError('Compose transform list size must be greater than 1')
    curr = transforms[-1]
    for nxt in reversed(transforms[:-1]):
        curr = curr + transform(nxt, curr, interp_method)
    return curr


def spatial_transformer(vol, trf, interp_method='linear', fill_value=None):
    return torch.stack([transform(v, t, interp_method, fill_value=fill_value)
                        for v, t in zip(vol, trf)], 0)


def vec_int(svf, int_steps=7):
    return torch.stack([integrate_vec(v, int_steps) for v in svf], 0)


def headline_pipeline(svf_half, image, int_steps=7):
    """BASELINE headline: VecInt(int_steps) @half-res -> RescaleTransform(2) -> linear warp."""
    flow = vec_int(svf_half, int_steps)
    flow = rescale_dense_transform(flow, 2)
    return spatial_transformer(image, flow), flow


# --------------------------------------------------------------------------------------
# losses adjacent to the warp (train_synthmorph.py:301-306) -- parity unpinned ([UR] voxelmorph.losses)
# --------------------------------------------------------------------------------------
def dice_loss(y_true, y_pred):
    """vxm.losses.Dice().loss: top = 2 sum(t p), bottom = sum(t + p) over the volume axes,
    -mean(divide_no_nan(top, bottom)); the same structure as the in-repo losses.py:57-68."""
    nd = y_pred.dim() - 2
    axes = tuple(range(1, nd + 1))
    top = 2 * (y_true * y_pred).sum(axes)
    bottom = (y_true + y_pred).sum(axes)
    safe = torch.where(bottom != 0, bottom, torch.ones_like(bottom))
    dice = torch.where(bottom != 0, top / safe, torch.zeros_like(top))
    return -dice.mean()


def grad_l2_loss(flow, loss_mult=1.0):
    """vxm.losses.Grad('l2', loss_mult).loss(None, flow): forward differences along every spatial axis,
    squared, mean over all elements per batch item, averaged over the axes; returns [B]."""
    nd = flow.dim() - 2
    terms = []
    for a in range(1, nd + 1):
        d = flow.narrow(a, 1, flow.shape[a] - 1) - flow.narrow(a, 0, flow.shape[a] - 1)
        terms.append((d * d).reshape(flow.shape[0], -1).mean(-1))
    return sum(terms) / nd * loss_mult


def dice_loss_zeropad(y_true, y_pred):
    """losses.py:11-69 as documented (the shipped function raises at :32 before reaching this code): batch
    item 0 only, voxels where channel 0 of either map is >= 1 are zeroed in every channel, Dice over the
    channels 1.., -mean(divide_no_nan(top, bottom))."""
    is0 = (y_true[0, ..., 0] >= 1) | (y_pred[0, ..., 0] >= 1)                  # :38-42
    keep = (~is0).to(y_pred.dtype)
    t = torch.stack([y_true[0, ..., i] * keep for i in range(y_pred.shape[-1])])   # :50-56 (channel-major stack)
    p = torch.stack([y_pred[0, ..., i] * keep for i in range(y_pred.shape[-1])])
    top = 2 * (t * p).sum((1, 2, 3))                                              # :58-59 (vol_axes on the stacked array)
    bottom = (t + p).sum((1, 2, 3))
    top, bottom = top[1:], bottom[1:]                                             # :62-63
    safe = torch.where(bottom != 0, bottom, torch.ones_like(bottom))
    return -torch.where(bottom != 0, top / safe, torch.zeros_like(top)).mean()    # :65-68
