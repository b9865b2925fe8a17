#!/usr/bin/env python
"""Drop-in for the reference's eval_reg_with_jacobian.py (same flags, same CSV columns, same output
volume), with the Jacobian determinant map computed on the GPU by libdfm (dfm_jacdet).

Reference behaviour kept (eval_reg_with_jacobian.py:46-108): the field is a 5-D NIfTI
(H, W, D, 1, 3); determinants are produced for the interior [2:-2]^3; a voxel counts as folded when
det < 0 (zeros are not counted); the determinant volume is written as (H-4, W-4, D-4, 1) float64 with
the field's affine; one CSV row per call, header written when the file is new or --append 0.
"""
import argparse
import csv
import datetime
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_registration_b200 as mrb            # noqa: E402
from multimodal_registration_b200 import _nifti        # noqa: E402

HEADER = ['Timestamp', 'Subject', 'Percentage_negative_detJa[%]', 'Median_detJa', 'Mean_detJa', 'Std_detJa',
          'N_total_voxels', 'N_voxels_negatives_detJa']


def jacobian_summary(ddf):
    """ddf: (H, W, D, 1, 3) array.  Returns (det volume (H-4, W-4, D-4, 1) float64, summary dict)."""
    if ddf.ndim != 5 or ddf.shape[3] != 1 or ddf.shape[4] != 3:
        raise ValueError('expected a deformation field of shape (H, W, D, 1, 3), got %s' % (ddf.shape,))
    dtype = torch.float64 if ddf.dtype == np.float64 else torch.float32
    field = torch.from_numpy(np.ascontiguousarray(ddf[:, :, :, 0, :])).to(dtype).cuda()[None]
    det, stats = mrb.ops.jacobian_determinant(field, out_dtype=torch.float64)
    n_neg, s, s2, n = [float(v) for v in stats[0].tolist()]
    flat = det.reshape(-1).sort().values                    # median only (a CSV statistic, not the hot path)
    m = flat.numel()
    median = 0.5 * (flat[(m - 1) // 2].item() + flat[m // 2].item())
    mean = s / n
    summary = {
        'percentage_negative_detJa': 100 * n_neg / n,
        'median_detJa': median,
        'mean_detJa': mean,
        'std_detJa': float(np.sqrt(max(s2 / n - mean * mean, 0.0))),
        'n_total_detJa': int(n),
        'n_negatives_detJa': int(n_neg),
    }
    return det[0].cpu().numpy()[..., None], summary


def main(argv=None):
    p = argparse.ArgumentParser(formatter_class=argparse.RawDescriptionHelpFormatter,
                                description='Evaluate the registration of two volumes using the deformation field')
    p.add_argument('--def-field-path', required=True, help='path to the deformation field (NIfTI, (H, W, D, 1, 3))')
    p.add_argument('--sub-id', required=True, help='id of the subject')
    p.add_argument('--out-file', required=False, default='jacobian_det.csv',
                   help='path to csv summarizing the results obtained')
    p.add_argument('--out-im-path', required=False, default='detJa.nii.gz',
                   help='path to output the volume representative of the determinant of the Jacobian')
    p.add_argument('--append', type=int, required=False, default=1, choices=[0, 1],
                   help='Append results as a new line in the output csv file instead of overwriting it.')
    arg = p.parse_args(argv)

    path = arg.def_field_path if len(arg.def_field_path.split('.')) > 1 else arg.def_field_path + '.nii.gz'
    ddf, affine = _nifti.load_nifti(path)
    det_vol, res = jacobian_summary(np.asarray(ddf))
    _nifti.save_nifti(det_vol, arg.out_im_path, affine)

    if not arg.append or not os.path.isfile(arg.out_file):
        with open(arg.out_file, 'w') as f:
            csv.DictWriter(f, fieldnames=HEADER).writeheader()
    with open(arg.out_file, 'a') as f:
        row = [datetime.datetime.now().strftime('%Y-%m-%d %H:%M:%S'), arg.sub_id] + [str(v) for v in res.values()]
        csv.writer(f, delimiter=',').writerow(row)
    return 0


if __name__ == '__main__':
    sys.exit(main())
