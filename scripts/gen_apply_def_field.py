#!/usr/bin/env python
"""Drop-in for the reference's gen_apply_def_field.py (same flags and outputs): draw a Perlin-like
deformation field, save it, and warp the input volume with it on the GPU (dfm_warp_fwd through the
voxelmorph mirror).  As in the reference (gen_apply_def_field.py:59-76) the field is applied directly
as a displacement and round-trips through a NIfTI file between generation and use.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_registration_b200 as mrb            # noqa: E402
from multimodal_registration_b200 import _nifti        # noqa: E402

vxm, ne = mrb.voxelmorph, mrb.neurite


def main(argv=None):
    p = argparse.ArgumentParser(formatter_class=argparse.RawDescriptionHelpFormatter,
                                description='Deform an image with the generated deformation field')
    p.add_argument('--im-path', required=True, help='path to the volume to deform')
    p.add_argument('--res-dir', required=False, default='res', help='results output directory (default: res)')
    p.add_argument('--out-im-name', default='moved_im', help='path where the moved volume will be saved')
    p.add_argument('--out-def-name', default='deformation_field', help='path where the def. field will be saved')
    p.add_argument('--def-scales', type=int, nargs='+', default=[16, 32, 64],
                   help='list of relative resolutions at which noise is sampled normally (default: 16 32 64)')
    p.add_argument('--def-max-std', type=int, default=3,
                   help='max std for the gaussian dist of noise in label maps generation (def field) (default: 3)')
    p.add_argument('--interp', default='linear', help='interpolation method linear/nearest (default: linear)')
    p.add_argument('--seed', type=int, default=None, help='(extension) seed of the noise generator')
    arg = p.parse_args(argv)

    vol, affine = _nifti.load_nifti(arg.im_path)
    shape = vol.shape[:3]
    os.makedirs(arg.res_dir, exist_ok=True)

    seeds = {'noise': arg.seed} if arg.seed is not None else None
    field = ne.utils.augment.draw_perlin(out_shape=(*shape, 1, 3), scales=arg.def_scales,
                                         max_std=arg.def_max_std, seeds=seeds)
    def_path = os.path.join(arg.res_dir, arg.out_def_name + '.nii.gz')
    _nifti.save_nifti(field[..., 0, :].cpu().numpy(), def_path, affine)

    moving = vxm.py.utils.load_volfile(arg.im_path, add_batch_axis=True, add_feat_axis=True)
    deform, _ = vxm.py.utils.load_volfile(def_path, add_batch_axis=True, ret_affine=True)
    moved = vxm.networks.Transform(moving.shape[1:-1], interp_method=arg.interp,
                                   nb_feats=moving.shape[-1]).predict([moving, deform])
    vxm.py.utils.save_volfile(moved.squeeze(), os.path.join(arg.res_dir, arg.out_im_name + '.nii.gz'), affine)
    return 0


if __name__ == '__main__':
    sys.exit(main())
