"""NumPy fp32 restatement of the voxelmorph / neurite deformation ops -- TEST ORACLE ONLY.

PARITY UNPINNED (see ``oracle/__init__.py``): the reference calls these functions from
un-vendored packages (reference ``README.md:35-37``: voxelmorph@52dd120f, neurite@c7bb05d5);
this file restates their published algorithm, keeping TensorFlow's
op order: every multiply and add below is one separately rounded fp32 operation (TF's
Eigen element-wise kernels do not contract to FMA), corners are visited in
``itertools.product([0, 1], repeat=N)`` order and the weight product is left-to-right.

Reference call sites each function stands in for (files under /root/reference):
  interpn                  -- neurite.utils.interpn, reached from every site below
  resize                   -- neurite.utils.resize, via rescale_dense_transform / draw_perlin
  transform                -- vxm.utils.transform: train_synthmorph.py:67 (channel-wise),
                              and inside SpatialTransformer (train_synthmorph.py:298,
                              gen_apply_def_field.py:74-76, 3d_reg.py:331-334,377-380)
  integrate_vec            -- vxm.layers.VecInt inside VxmDense (3d_reg.py:305,
                              bids_two_steps_registration.py:311,314, train_synthmorph.py:296)
  rescale_dense_transform  -- 3d_reg.py:394, bids_registration.py:398,
                              bids_two_steps_registration.py:515
  compose                  -- bids_two_steps_registration.py:324,346,369,484
"""
import itertools

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def linspace_tf(start, stop, num):
    """``tf.linspace`` (TF 2.x ``linspace_nd``) in fp32: ``[start, start + delta*k ..., stop]``.

    delta = (stop - start) / (num - 1) is one fp32 division; interior points are
    ``start + delta * k`` (fp32 multiply, then fp32 add); first and last entries are the
    exact end points.  ``num == 1`` yields ``[start]``.   SURVEY.md Appendix A.2.
    """
    start = F32(start)
    stop = F32(stop)
    num = int(num)
    if num <= 0:
        return np.zeros((0,), F32)
    if num == 1:
        return np.array([start], F32)
    delta = F32(F32(stop - start) / F32(num - 1))
    k = np.arange(1, num - 1, dtype=np.int64).astype(F32)
    inner = (start + (delta * k).astype(F32)).astype(F32)
    return np.concatenate([[start], inner, [stop]]).astype(F32)


def volshape_to_meshgrid(volshape, indexing='ij'):
    """neurite ``volshape_to_meshgrid``: integer ranges, meshgrid'ed (returned as int arrays)."""
    if indexing != 'ij':
        raise ValueError('oracle restates ij indexing only (the reference uses the default)')
    return np.meshgrid(*[np.arange(int(d)) for d in volshape], indexing='ij')


def _prod_n(lst):
    prod = lst[0]
    for p in lst[1:]:
        prod = (prod * p).astype(F32)
    return prod


def _sub2ind(siz, subs):
    # C-order flat index (neurite sub2ind2d)
    k = np.cumprod(np.asarray(siz[::-1], dtype=np.int64))
    ndx = subs[-1].astype(np.int64)
    for i, v in enumerate(subs[:-1][::-1]):
        ndx = ndx + v.astype(np.int64) * k[i]
    return ndx


# --------------------------------------------------------------------------------------
# neurite.utils.interpn
# --------------------------------------------------------------------------------------
def interpn(vol, loc, interp_method='linear', fill_value=None):
    """N-D gather interpolation with edge clamp.

    vol: [*vol_shape] or [*vol_shape, C];  loc: list of N arrays or [*out_shape, N].
    Returns [*out_shape, C] (always with a trailing feature axis, like neurite).
    """
    if isinstance(loc, (list, tuple)):
        loc = np.stack([np.asarray(l) for l in loc], -1)
    loc = np.asarray(loc)
    nb_dims = loc.shape[-1]
    vol = np.asarray(vol)
    if nb_dims != vol.ndim - 1 and nb_dims != vol.ndim:
        raise ValueError("Number of loc Tensors %d does not match volume dimension %d"
                         % (nb_dims, vol.ndim - 1))
    if vol.ndim == nb_dims:
        vol = vol[..., None]
    loc = loc.astype(F32)
    volshape = vol.shape[:-1]
    max_loc = [d - 1 for d in volshape]
    vol_flat = vol.reshape(-1, vol.shape[-1])

    if interp_method == 'linear':
        if vol.dtype != F32:
            vol_flat = vol_flat.astype(F32)
        loc0 = np.floor(loc)
        clipped = [np.clip(loc[..., d], F32(0), F32(max_loc[d])) for d in range(nb_dims)]
        loc0lst = [np.clip(loc0[..., d], F32(0), F32(max_loc[d])) for d in range(nb_dims)]
        loc1 = [np.clip((loc0lst[d] + F32(1)).astype(F32), F32(0), F32(max_loc[d]))
                for d in range(nb_dims)]
        locs = [[f.astype(np.int32) for f in loc0lst], [f.astype(np.int32) for f in loc1]]
        diff_loc1 = [(loc1[d] - clipped[d]).astype(F32) for d in range(nb_dims)]
        diff_loc0 = [(F32(1) - d).astype(F32) for d in diff_loc1]
        weights_loc = [diff_loc1, diff_loc0]

        interp_vol = np.zeros(loc.shape[:-1] + (vol.shape[-1],), F32)
        for c in itertools.product([0, 1], repeat=nb_dims):
            subs = [locs[c[d]][d] for d in range(nb_dims)]
            idx = _sub2ind(volshape, subs)
            vol_val = vol_flat[idx]
            wt = _prod_n([weights_loc[c[d]][d] for d in range(nb_dims)])
            interp_vol = (interp_vol + (wt[..., None] * vol_val).astype(F32)).astype(F32)
    elif interp_method == 'nearest':
        # tf.round is round-half-to-even, applied to the UNCLIPPED location
        roundloc = np.rint(loc).astype(np.int32)
        roundloc = [np.clip(roundloc[..., d], 0, max_loc[d]) for d in range(nb_dims)]
        idx = _sub2ind(volshape, roundloc)
        interp_vol = vol_flat[idx]
    else:
        raise ValueError("interp_method must be 'linear' or 'nearest', got %r" % (interp_method,))

    if fill_value is not None:
        out_type = interp_vol.dtype
        fv = np.asarray(fill_value).astype(out_type)
        below = [loc[..., d] < 0 for d in range(nb_dims)]
        above = [loc[..., d] > max_loc[d] for d in range(nb_dims)]
        oob = np.any(np.stack(below + above, -1), -1, keepdims=True)
        interp_vol = (interp_vol * np.logical_not(oob).astype(out_type)).astype(out_type)
        interp_vol = (interp_vol + (oob.astype(out_type) * fv).astype(out_type)).astype(out_type)
    return interp_vol


# --------------------------------------------------------------------------------------
# neurite.utils.resize  (alias zoom)
# --------------------------------------------------------------------------------------
def resize(vol, zoom_factor, interp_method='linear'):
    """Corner-aligned resample onto ``linspace(0, n_in-1, int(n_in*zoom))`` (Appendix A.2)."""
    vol = np.asarray(vol)
    if isinstance(zoom_factor, (list, tuple)):
        ndims = len(zoom_factor)
        vol_shape = vol.shape[:ndims]
        zoom = list(zoom_factor)
    else:
        vol_shape = vol.shape[:-1]
        ndims = len(vol_shape)
        zoom = [zoom_factor] * ndims
    new_shape = [int(vol_shape[d] * zoom[d]) for d in range(ndims)]
    lin = [linspace_tf(0., vol_shape[d] - 1., new_shape[d]) for d in range(ndims)]
    grid = np.meshgrid(*lin, indexing='ij')
    return interpn(vol, grid, interp_method=interp_method)


# --------------------------------------------------------------------------------------
# voxelmorph.utils
# --------------------------------------------------------------------------------------
def transform(vol, loc_shift, interp_method='linear', indexing='ij', fill_value=None):
    """``out[x] = interp(vol, x + loc_shift[x])``; channel-wise when the shift carries a
    channel axis ``[*vol_shape, C, D]`` (Appendix A.3)."""
    vol = np.asarray(vol)
    loc_shift = np.asarray(loc_shift)
    loc_volshape = loc_shift.shape[:-1]
    nb_dims = vol.ndim - 1
    is_channelwise = len(loc_volshape) == nb_dims + 1
    if loc_shift.shape[-1] != nb_dims:
        raise ValueError("Dimension check failed for ne.utils.transform(): vol has %d spatial dims, "
                         "loc_shift has %d components" % (nb_dims, loc_shift.shape[-1]))
    mesh = volshape_to_meshgrid(loc_volshape, indexing=indexing)
    shift = loc_shift.astype(F32)
    loc = [(mesh[d].astype(F32) + shift[..., d]).astype(F32) for d in range(nb_dims)]
    if is_channelwise:
        loc.append(mesh[-1].astype(F32))
    out = interpn(vol, loc, interp_method=interp_method, fill_value=fill_value)
    if is_channelwise:
        out = out[..., 0]
    return out


def integrate_vec(vec, nb_steps):
    """Scaling and squaring: ``v /= 2**n``; n times ``v += transform(v, v)`` (Appendix A.4)."""
    vec = np.asarray(vec).astype(F32)
    if nb_steps < 0:
        raise ValueError('nb_steps should be >= 0, found: %d' % nb_steps)
    vec = (vec / F32(2 ** nb_steps)).astype(F32)
    for _ in range(nb_steps):
        vec = (vec + transform(vec, vec)).astype(F32)
    return vec


def rescale_dense_transform(trf, factor, interp_method='linear'):
    """Unbatched ``[*shape, D]`` or batched ``[B, *shape, D]`` (mapped item by item) (A.5)."""
    trf = np.asarray(trf)
    if trf.ndim > trf.shape[-1] + 1:
        return np.stack([rescale_dense_transform(t, factor, interp_method) for t in trf], 0)
    trf = trf.astype(F32)
    if factor < 1:
        trf = resize(trf, factor, interp_method=interp_method)
        return (trf * F32(factor)).astype(F32)
    trf = (trf * F32(factor)).astype(F32)
    return resize(trf, factor, interp_method=interp_method)


def compose(transforms, interp_method='linear', shift_center=True, indexing='ij'):
    """Right fold ``curr = curr + transform(nxt, curr)`` for dense shifts (Appendix A.6)."""
    if indexing != 'ij':
        raise ValueError('Compose transform only supports ij indexing')
    if len(transforms) < 2:
        raise ValueError('Compose transform list size must be greater than 1')
    curr = np.asarray(transforms[-1]).astype(F32)
    for nxt in reversed(transforms[:-1]):
        nxt = np.asarray(nxt).astype(F32)
        curr = (curr + transform(nxt, curr, interp_method=interp_method)).astype(F32)
    return curr


# --------------------------------------------------------------------------------------
# layers (batched; tf.map_fn over the batch axis)  -- Appendix A.7 / A.8
# --------------------------------------------------------------------------------------
def spatial_transformer(vol, trf, interp_method='linear', fill_value=None, single_transform=False):
    vol = np.asarray(vol)
    trf = np.asarray(trf)
    if single_transform:
        return np.stack([transform(v, trf[0], interp_method, fill_value=fill_value) for v in vol], 0)
    return np.stack([transform(v, t, interp_method, fill_value=fill_value)
                     for v, t in zip(vol, trf)], 0)


def vec_int(svf, int_steps=7):
    return np.stack([integrate_vec(v, int_steps) for v in np.asarray(svf)], 0)


def rescale_transform(trf, zoom_factor, interp_method='linear'):
    return rescale_dense_transform(np.asarray(trf), zoom_factor, interp_method)


def transform_model(scan, trf, interp_method='linear', rescale=None, fill_value=None):
    """vxm.networks.Transform(...).predict([scan, trf]) for dense fields (Appendix A.8)."""
    if rescale is not None:
        trf = rescale_transform(trf, rescale, 'linear')   # RescaleTransform(rescale): default interp
    return spatial_transformer(scan, trf, interp_method, fill_value)
