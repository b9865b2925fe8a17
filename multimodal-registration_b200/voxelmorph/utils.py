"""vxm.utils mirror: transform, compose, rescale_dense_transform, integrate_vec.

Signatures follow voxelmorph; tensors are channels-last like the
reference.  Inputs may be numpy arrays, CPU tensors or CUDA tensors; results are CUDA tensors
(the analogue of TF eager tensors) -- use ``to_numpy`` (the analogue of ``K.eval``) to fetch.
"""
import torch

from .. import _host, ops


def to_numpy(t):
    """K.eval analogue (bids_two_steps_registration.py:325)."""
    return _host.to_host(t)


def is_affine_shape(shape):
    return len(shape) == 1 or (len(shape) == 2 and shape[0] + 1 == shape[1])


def transform(vol, loc_shift, interp_method='linear', indexing='ij', fill_value=None):
    """Unbatched ``vol [X, Y, Z, C]`` warped by ``loc_shift [X', Y', Z', 3]`` (output on the
    shift's grid), or channel-wise when ``loc_shift`` is ``[X, Y, Z, C, 3]``
    (train_synthmorph.py:67)."""
    if indexing != 'ij':
        raise ValueError("transform: only indexing='ij' is supported (the reference default)")
    vol = _host.to_device(vol)
    loc_shift = _host.to_device(loc_shift, torch.float32)
    nb_dims = vol.dim() - 1
    if nb_dims != 3:
        raise NotImplementedError('transform: only 3-D volumes are supported, got %d-D' % nb_dims)
    if loc_shift.shape[-1] != nb_dims:
        raise ValueError('Dimension check failed for ne.utils.transform(): {}D volume (shape {}) '
                         'called with {}D transform'.format(nb_dims, tuple(vol.shape[:-1]), loc_shift.shape[-1]))
    if loc_shift.dim() == nb_dims + 2:        # channel-wise
        return ops.warp_channelwise(vol[None], loc_shift[None], interp_method, fill_value)[0]
    return ops.warp(vol[None], loc_shift[None], interp_method, fill_value)[0]


def integrate_vec(vec, time_dep=False, method='ss', **kwargs):
    """Scaling-and-squaring integration of an unbatched SVF ``[X, Y, Z, 3]``."""
    if method not in ('ss', 'scaling_and_squaring'):
        raise NotImplementedError("integrate_vec: only method='ss' is implemented (the only one "
                                  "the reference uses); got %r" % (method,))
    if time_dep:
        raise NotImplementedError('integrate_vec: time-dependent fields are not used by the reference')
    nb_steps = kwargs['nb_steps']
    if nb_steps < 0:
        raise ValueError('nb_steps should be >= 0, found: %d' % nb_steps)
    vec = _host.to_device(vec, torch.float32)
    return ops.vecint(vec[None], nb_steps)[0]


def rescale_dense_transform(transform, factor, interp_method='linear'):
    """Unbatched ``[X, Y, Z, 3]`` or batched ``[B, X, Y, Z, 3]`` (3d_reg.py:394)."""
    trf = _host.to_device(transform, torch.float32)
    if trf.dim() > trf.shape[-1] + 1:
        return ops.rescale_dense_transform(trf, factor, interp_method)
    return ops.rescale_dense_transform(trf[None], factor, interp_method)[0]


def compose(transforms, interp_method='linear', shift_center=True, indexing='ij'):
    """compose([A, B]) = B + A o (id + B) for unbatched dense shifts
    (bids_two_steps_registration.py:324)."""
    if indexing != 'ij':
        raise ValueError('Compose transform only supports ij indexing')
    if len(transforms) < 2:
        raise ValueError('Compose transform list size must be greater than 1')
    dev = []
    for t in transforms:
        t = _host.to_device(t, torch.float32)
        if is_affine_shape(t.shape):
            raise NotImplementedError('compose: affine transforms are not on the reference hot path')
        dev.append(t[None])
    return ops.compose(dev, interp_method)[0]
