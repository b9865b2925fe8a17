"""SCT warp convention at the edge of the path.

The reference turns the model's displacement field into a file `sct_apply_transfo` can consume
(3d_reg.py:390-422, bids_registration.py:394-426): rescale to the fixed image's resolution, add a
singleton time axis, re-order and sign-flip the vector components from the fixed image's orientation to
"RAI", and mark the NIfTI as a vector field (`intent_code = 1007`).  The rescale is the x2 kernel of the
hot path (`ops.rescale_dense_transform`); the permutation / sign is a per-component epilogue done on the
device before the one device-to-host copy.
"""
import numpy as np
import torch

from . import _nifti, ops

ORIENTATION_CONV = 'RAI'                                   # 3d_reg.py:399
_OPPOSITE = {'L': 'R', 'R': 'L', 'A': 'P', 'P': 'A', 'I': 'S', 'S': 'I'}


def rai_permutation(axcodes):
    """(perm, inversion) of 3d_reg.py:403-411: component i of the exported vector is
    ``inversion[i] * v[perm[i]]``, where ``axcodes`` are ``aff2axcodes(-fixed_affine)``."""
    axcodes = list(axcodes)
    perm, inversion = [0, 1, 2], [1, 1, 1]
    for i, ch in enumerate(ORIENTATION_CONV):
        if ch in axcodes:
            perm[i] = axcodes.index(ch)
        else:
            perm[i] = axcodes.index(_OPPOSITE[ch])
            inversion[i] = -1
    return perm, inversion


def to_sct_warp(warp, fixed_affine, scale=1):
    """warp [X, Y, Z, 3] (numpy or tensor, voxel displacements at the model's field resolution) ->
    float32 numpy [X*scale, Y*scale, Z*scale, 1, 3] in the SCT convention (3d_reg.py:393-418)."""
    t = warp if isinstance(warp, torch.Tensor) else torch.as_tensor(np.asarray(warp, dtype=np.float32))
    if t.dim() != 4 or t.shape[-1] != 3:
        raise ValueError('warp must be [X, Y, Z, 3], got %s' % (tuple(t.shape),))
    t = t.float().cuda()[None]
    t = ops.rescale_dense_transform(t, scale)              # :394 (identity resample when scale == 1, like the reference)
    perm, inversion = rai_permutation(_nifti.aff2axcodes(-np.asarray(fixed_affine, dtype=np.float64)))
    t = ops.to_layout(t, 'cl')[0]
    sign = torch.tensor(inversion, dtype=t.dtype, device=t.device)
    t = t[..., perm] * sign                                # :414-416
    return t.unsqueeze(3).cpu().numpy()                    # :412 time axis


def save_sct_warp(path, warp, fixed_affine, scale=1):
    """Write the SCT-convention warp as a NIfTI-1 vector field (intent_code 1007, 3d_reg.py:417-420)."""
    data = to_sct_warp(warp, fixed_affine, scale)
    _nifti.save_nifti(data, path, affine=np.asarray(fixed_affine, dtype=np.float64), intent_code=1007)
    return data
