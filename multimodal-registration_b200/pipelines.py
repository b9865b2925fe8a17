"""Device-resident tails of the reference's registration scripts (BASELINE.json config 4).

``bids_two_steps_registration.py:register`` (:316-355 whole volume, :357-500 sub-volumes, :504-546 export)
chains two ``VxmDense`` models, ``vxm.utils.compose``, ``Transform(rescale=scale)`` and
``rescale_dense_transform`` -- with a Keras ``predict`` (device -> host), a NIfTI file written and re-read, and
a fresh model object between every pair of steps.  Here the same chain runs on device tensors from the first
kernel to the exported warp: no host hop, no file round trip.  The U-Nets are outside the hot path: the flows
their flow convolutions emit come in as tensors (``flow1``) or as a callable (``flow2``, which needs the first
step's moved image).

Names follow the reference: ``warp_first_reg`` / ``warp_second_reg`` are the models' second outputs (half
resolution under ``int_res = 2``), ``scale`` is 1 or 2 exactly as the script derives it (:323).
"""
import numpy as np
import torch

from . import _host, ops, sct_warp
from .voxelmorph import networks


def _as_flow(flow, *args):
    return flow(*args) if callable(flow) else flow


def _scale_of(warp, inshape):
    # :323  scale = 1 if warp_first_reg[0, ...].shape[0] == model_in_shape[0] else 2
    return 1 if int(warp.shape[1]) == int(inshape[0]) else 2


@torch.no_grad()
def two_steps_tail(moving, fixed, model1, model2, flow1, flow2, warp_interp='linear', moving_proc=None):
    """Whole-volume branch of ``register`` (bids_two_steps_registration.py:316-355).

    moving, fixed: [B, X, Y, Z, 1] (host or device); model1 / model2: ``voxelmorph.networks.VxmDense`` tails;
    flow1: the first U-Net's flow [B, x, y, z, 3]; flow2: tensor, or callable ``(moved_first, fixed) -> flow``.
    moving_proc: the image ``Transform`` is applied to in the non-linear branch (the reference re-loads
    ``*_proc.nii.gz``, :332 -- the same volume as ``moving`` unless the caller says otherwise).

    Returns a dict of device tensors: ``moved`` [B, X, Y, Z, 1], ``warp`` (composed, at the models' field
    resolution), ``scale``, ``moved_first_reg``, ``warp_first_reg``, ``warp_second_reg``.
    """
    moving = _host.to_device(moving, torch.float32, tag='moving')
    fixed = _host.to_device(fixed, torch.float32, tag='fixed')
    inshape = model1.inshape
    if warp_interp == 'linear':
        moved_first, warp_first = model1.deform([moving, _host.to_device(flow1, torch.float32, tag='flow1')], keep_pos_flow=False)   # :318-319
        moved, warp_second = model2.deform([moved_first, _as_flow(flow2, moved_first, fixed)], keep_pos_flow=False)     # :320-321
        scale = _scale_of(warp_first, inshape)
        warp = ops.compose([warp_first, warp_second])                                                                  # :324
    else:
        _, warp_first = model1.deform([moving, _host.to_device(flow1, torch.float32, tag='flow1')], keep_pos_flow=False)   # :328-329
        scale = _scale_of(warp_first, inshape)
        src = moving if moving_proc is None else _host.to_device(moving_proc, torch.float32, tag='moving_proc')
        tr = networks.Transform(src.shape[1:-1], interp_method=warp_interp, rescale=scale, nb_feats=src.shape[-1])
        moved_first = tr([src, warp_first])                                                                            # :338-341
        _, warp_second = model2.deform([moved_first, _as_flow(flow2, moved_first, fixed)], keep_pos_flow=False)       # :343-344
        warp = ops.compose([warp_first, warp_second])                                                                  # :346
        moved = tr([src, warp])                                                                                        # :354-355
    return dict(moved=moved, warp=warp, scale=scale, moved_first_reg=moved_first, warp_first_reg=warp_first,
                warp_second_reg=warp_second)


def _halve(coords, shape, half):
    # :373-386: coordinates and shapes of a half-resolution field
    if not half:
        return [tuple(int(v) for v in c) for c in coords], tuple(int(d) for d in shape)
    return [tuple(int(v) // 2 for v in c) for c in coords], tuple(int(d) // 2 for d in shape)


@torch.no_grad()
def two_steps_tail_subvol(moving, lst_subvol_mov, lst_subvol_fx, lst_coords_subvol, model1, model2, flows1, flows2,
                          warp_interp='linear'):
    """Sub-volume branch (bids_two_steps_registration.py:357-400, linear): per sub-volume two model tails and a
    ``compose``, then ``get_def_field_from_subvol`` (pyramid-weighted stitching, :226-271) and one
    ``Transform(rescale=scale)`` of the whole moving image.  flows1[k]: flow of sub-volume k; flows2[k]: tensor or
    callable ``(moved_first, fixed_subvol) -> flow``.  Returns ``moved``, ``warp`` (stitched), ``scale``."""
    if warp_interp != 'linear':
        raise NotImplementedError('two_steps_tail_subvol: the nearest branch (:402-500) re-runs the preprocessing '
                                  'between the two registrations; run two_steps_tail per stage and ops.stitch_subvolumes')
    moving = _host.to_device(moving, torch.float32, tag='moving')
    fields = []
    for k, (mov, fx) in enumerate(zip(lst_subvol_mov, lst_subvol_fx)):
        mov = _host.to_device(mov, torch.float32, tag='subvol_mov')
        fx = _host.to_device(fx, torch.float32, tag='subvol_fx')
        moved_first, w1 = model1.deform([mov, _host.to_device(flows1[k], torch.float32, tag='flow1')], keep_pos_flow=False)   # :361-362
        _, w2 = model2.deform([moved_first, _as_flow(flows2[k], moved_first, fx)], keep_pos_flow=False)               # :363-364
        fields.append(ops.to_layout(ops.compose([w1, w2]), 'cl')[0])                                                   # :365-367
    half = int(fields[0].shape[0]) != int(model1.inshape[0])                                                          # :369
    scale = 2 if half else 1
    coords, im_shape = _halve(lst_coords_subvol, moving.shape[1:4], half)
    tile_shape = tuple(int(d) // 2 for d in model1.inshape) if half else tuple(model1.inshape)
    warp = ops.stitch_subvolumes(tile_shape, im_shape, coords, torch.stack(fields, 0), out_dtype=torch.float32)        # :388
    tr = networks.Transform(moving.shape[1:-1], interp_method=warp_interp, rescale=scale, nb_feats=moving.shape[-1])
    moved = tr([moving, warp[None]])                                                                                  # :398-399
    return dict(moved=moved, warp=warp[None], scale=scale)


@torch.no_grad()
def export_sct_warp(warp, scale, fixed_affine, path=None):
    """:504-546: ``rescale_dense_transform(warp, scale)``, time axis, RAI permutation / signs, intent 1007.
    warp [1, x, y, z, 3] or [x, y, z, 3]; returns the float32 array [X, Y, Z, 1, 3] (and writes ``path``)."""
    w = warp[0] if warp.dim() == 5 else warp
    w = ops.to_layout(w[None], 'cl')[0]
    if path is not None:
        return sct_warp.save_sct_warp(path, w, fixed_affine, scale)
    return sct_warp.to_sct_warp(w, fixed_affine, scale)
