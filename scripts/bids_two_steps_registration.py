#!/usr/bin/env python
"""Drop-in for the POST-NETWORK part of the reference's bids_two_steps_registration.py (config 4): the cascaded
two-step registration tail -- two ``VxmDense`` deformation tails, ``vxm.utils.compose``, the final
``Transform`` (linear or nearest), ``rescale_dense_transform`` and the SCT warp file -- resident on the GPU from
the first kernel to the exported warp (multimodal_registration_b200.pipelines).

Same flags as the reference (bids_two_steps_registration.py:562-597), same config file (config/config_inference.json),
same output file names (``<mov>_proc_reg_to_<contrast>.nii.gz`` :503, ``<mov>_proc_field_to_<contrast>.nii.gz`` :546).
What is outside the hot path is taken as input instead of being recomputed:
  * the U-Nets: ``--model1-path`` / ``--model2-path`` name what each network's flow convolution emits for this
    subject -- a ``.npy`` / ``.nii(.gz)`` array ``[x, y, z, 3]`` (the half-resolution SVF), or ``module:function``
    for a Python callable ``(source, target) -> flow`` (both are device tensors ``[1, X, Y, Z, 1]``);
  * the preprocessing (nilearn resampling, padding, intensity scaling, :100-167): ``--fx-img-path`` /
    ``--mov-img-path`` are the ``*_proc.nii.gz`` volumes the reference writes at :293-294, or any pair of volumes of
    one shape that is a multiple of 16;
  * the resampling back to the moving image's original grid (nilearn ``resample_img``, :508-511,548-551).
``--one-cpu-tf`` is accepted and ignored (it has no effect under TF2 either).
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_registration_b200 as mrb                                  # noqa: E402
from multimodal_registration_b200 import _nifti, ops, pipelines             # noqa: E402
from multimodal_registration_b200.voxelmorph import networks                # noqa: E402


def load_flow(spec):
    """``module:function`` -> callable; otherwise an array file -> tensor [1, x, y, z, 3]."""
    if ':' in spec and not os.path.exists(spec):
        mod, fn = spec.split(':', 1)
        return getattr(importlib.import_module(mod), fn)
    arr = np.load(spec) if spec.endswith('.npy') else _nifti.load_nifti(spec)[0]
    arr = np.asarray(arr, dtype=np.float32)
    arr = arr.reshape(arr.shape[:3] + (3,)) if arr.ndim == 5 else arr        # (x, y, z, 1, 3) vector NIfTI
    if arr.ndim != 4 or arr.shape[-1] != 3:
        raise ValueError('%s: expected a flow of shape [x, y, z, 3], got %s' % (spec, arr.shape))
    return torch.from_numpy(np.ascontiguousarray(arr))[None]


def register(model_inference_specs, flow1, flow2, fx_im_path, mov_im_path, fx_contrast='T1w'):
    """bids_two_steps_registration.py:274-551 from the model calls on (whole-volume branch)."""
    warp_interp = model_inference_specs['warp_interpolation']
    if warp_interp not in ['nearest', 'linear']:                                                 # :281-282
        warp_interp = 'linear'
    if model_inference_specs.get('use_subvol'):
        raise NotImplementedError('sub-volume registration needs one flow per tile: call '
                                  'pipelines.two_steps_tail_subvol(...) with the tiles of preprocess() (:357-400)')
    fixed, fixed_affine = _nifti.load_nifti(fx_im_path)
    moving, _ = _nifti.load_nifti(mov_im_path)
    fixed = np.asarray(fixed, dtype=np.float32).squeeze()
    moving = np.asarray(moving, dtype=np.float32).squeeze()
    if fixed.shape != moving.shape or fixed.ndim != 3:
        raise ValueError('fixed %s and moving %s must be preprocessed 3-D volumes of one shape' % (fixed.shape, moving.shape))
    # :287-288 cut the path at its first dot; here only the file name is cut (a dot in a directory name breaks the reference)
    mov_base = os.path.join(os.path.dirname(mov_im_path), os.path.basename(mov_im_path).split('.')[0])
    if mov_base.endswith('_proc'):                     # given the preprocessed file itself: the names below re-append _proc
        mov_base = mov_base[:-5]
    inshape = fixed.shape                                                                         # :302
    reg_args = dict(inshape=inshape, int_steps=model_inference_specs['int_steps'],
                    int_resolution=model_inference_specs['int_res'], svf_resolution=model_inference_specs['svf_res'])
    model1, model2 = networks.VxmDense(**reg_args), networks.VxmDense(**reg_args)                 # :311-315
    mv = torch.from_numpy(moving)[None, ..., None].cuda()
    fx = torch.from_numpy(fixed)[None, ..., None].cuda()
    f1 = flow1(mv, fx) if callable(flow1) else flow1.cuda()
    f2 = flow2 if callable(flow2) else flow2.cuda()
    res = pipelines.two_steps_tail(mv, fx, model1, model2, f1, f2, warp_interp=warp_interp)      # :316-355
    moved = ops.to_layout(res['moved'], 'cl')[0, ..., 0].cpu().numpy()
    moved_path = '%s_proc_reg_to_%s.nii.gz' % (mov_base, fx_contrast)
    _nifti.save_nifti(moved, moved_path, fixed_affine)                                            # :503
    warp_path = '%s_proc_field_to_%s.nii.gz' % (mov_base, fx_contrast)
    pipelines.export_sct_warp(res['warp'], res['scale'], fixed_affine, warp_path)                 # :513-546
    return moved_path, warp_path


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--model1-path', required=True, type=str,
                        help='first registration step: the flow its network emits ([x, y, z, 3] .npy / .nii.gz) or module:function')
    parser.add_argument('--model2-path', required=True, type=str,
                        help='second registration step: the flow its network emits, or module:function (source, target) -> flow')
    parser.add_argument('--config-path', required=True, type=str,
                        help='path to the config file with the inference models specificities')
    parser.add_argument('--fx-img-path', required=True, help='path to the (preprocessed) fixed image')
    parser.add_argument('--mov-img-path', required=True, help='path to the (preprocessed) moving image')
    parser.add_argument('--fx-img-contrast', required=False, default='T1w',
                        help='contrast of the fixed image: one of {T1w, T2w, T2star}')
    parser.add_argument('--one-cpu-tf', required=False, type=str, default='True', help='accepted for compatibility; no effect')
    args = parser.parse_args(argv)
    with open(args.config_path) as config_file:
        model_inference_specs = json.load(config_file)
    register(model_inference_specs, load_flow(args.model1_path), load_flow(args.model2_path),
             args.fx_img_path, args.mov_img_path, args.fx_img_contrast)
    return 0


if __name__ == '__main__':
    sys.exit(main())
