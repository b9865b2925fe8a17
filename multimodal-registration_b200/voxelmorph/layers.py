"""vxm.layers mirror: SpatialTransformer, VecInt, RescaleTransform.

Keras-style call convention: ``layer([vol, trf])`` / ``layer(trf)`` on batched channels-last
tensors.  The layers are ``torch.nn.Module``s (no parameters) and differentiable through
the CUDA backward kernels.
"""
import warnings

import torch

from .. import _host, ops


def _swap_xy(trf):
    # indexing='xy': swap the first two vector components (voxelmorph layers)
    idx = torch.tensor([1, 0, 2], device=trf.device)
    return trf.index_select(-1, idx)


class SpatialTransformer(torch.nn.Module):
    """out[b] = transform(vol[b], trf[b])  (train_synthmorph.py:298; inside Transform/VxmDense)."""

    def __init__(self, interp_method='linear', indexing='ij', single_transform=False,
                 fill_value=None, shift_center=True, **kwargs):
        super().__init__()
        if indexing not in ('ij', 'xy'):
            raise ValueError("indexing has to be 'ij' (matrix) or 'xy' (cartesian)")
        self.interp_method = interp_method
        self.indexing = indexing
        self.single_transform = single_transform
        self.fill_value = fill_value
        self.shift_center = shift_center
        self.name = kwargs.get('name')

    def forward(self, inputs):
        if len(inputs) != 2:
            raise ValueError('Spatial Transformer must be called on a list of length 2: '
                             'first argument is the image, second is the transform.')
        vol = _host.to_device(inputs[0])
        trf = _host.to_device(inputs[1], torch.float32)
        if trf.dim() in (2, 3):
            raise NotImplementedError('SpatialTransformer: affine transforms are not on the '
                                      'reference hot path (Transform(affine=False) everywhere)')
        if vol.dim() != 5 or trf.dim() != 5:
            raise NotImplementedError('SpatialTransformer: expected [B, X, Y, Z, C] and [B, X, Y, Z, 3]')
        if tuple(trf.shape[1:-1]) != tuple(vol.shape[1:-1]):
            warnings.warn('Dense transform shape %s does not match image shape %s.'
                          % (tuple(trf.shape[1:-1]), tuple(vol.shape[1:-1])))
        if self.indexing == 'xy':
            trf = _swap_xy(trf)
        if self.single_transform:
            trf = trf[:1].expand(vol.shape[0], -1, -1, -1, -1).contiguous()
        lab = getattr(vol, 'dfm_labels', None)
        if (lab is not None and self.interp_method == 'linear' and tuple(lab[0].shape) == tuple(vol.shape[:4])
                and int(lab[1]) == int(vol.shape[-1]) and not vol.requires_grad):
            # a one-hot map from ne.models.labels_to_image: warp it from its label map (same bits, 8 bytes per voxel gathered)
            return ops.warp_onehot(lab[0], trf, lab[1], self.fill_value)
        return ops.warp(vol, trf, self.interp_method, self.fill_value)


class VecInt(torch.nn.Module):
    """Scaling-and-squaring integration of a batched SVF (inside VxmDense / labels_to_image)."""

    def __init__(self, indexing='ij', method='ss', int_steps=7, out_time_pt=1,
                 ode_args=None, odeint_fn=None, **kwargs):
        super().__init__()
        if indexing not in ('ij', 'xy'):
            raise ValueError("indexing has to be 'ij' (matrix) or 'xy' (cartesian)")
        if method not in ('ss', 'scaling_and_squaring'):
            raise NotImplementedError("VecInt: only method='ss' is implemented (the only one the "
                                      "reference uses); got %r" % (method,))
        self.indexing = indexing
        self.method = method
        self.int_steps = int_steps
        self.name = kwargs.get('name')

    def forward(self, inputs):
        svf = inputs[0] if isinstance(inputs, (list, tuple)) else inputs
        svf = _host.to_device(svf, torch.float32)
        if self.indexing == 'xy':
            svf = _swap_xy(svf)
        return ops.vecint(svf, self.int_steps)


class RescaleTransform(torch.nn.Module):
    """Rescale a batched dense transform: resize the grid and scale the vectors."""

    def __init__(self, zoom_factor, interp_method='linear', **kwargs):
        super().__init__()
        self.zoom_factor = zoom_factor
        self.interp_method = interp_method
        self.name = kwargs.get('name')

    def forward(self, transform):
        if isinstance(transform, (list, tuple)):
            if len(transform) != 1:
                raise ValueError('RescaleTransform must be called on one tensor')
            transform = transform[0]
        trf = _host.to_device(transform, torch.float32)
        if trf.dim() in (2, 3):
            raise NotImplementedError('RescaleTransform: affine transforms are not on the reference hot path')
        return ops.rescale_dense_transform(trf, self.zoom_factor, self.interp_method)
