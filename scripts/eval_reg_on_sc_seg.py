#!/usr/bin/env python
"""Drop-in for the reference's eval_reg_on_sc_seg.py (same flags, same CSV columns, same exit codes): overlap
metrics between the spinal-cord segmentations of the fixed, moving and moved images, with the voxel pass on
the GPU (multimodal_registration_b200.metrics.overlap_metrics -> dfm_overlap_sums).

Reference behaviour kept (eval_reg_on_sc_seg.py:72-180): volumes read as float64; TP / FP / TN / FN against
the fixed segmentation (voxels == 1 / == 0); exit status 1 when 100 * Dice(fixed, moved) < --min-dice and
--last-eval is 0 (no CSV row is written then); otherwise one CSV row and exit status 0.
"""
import argparse
import csv
import datetime
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_registration_b200 as mrb            # noqa: E402
from multimodal_registration_b200 import _nifti, metrics        # noqa: E402

HEADER = ['Timestamp', 'Subject', 'Dice_before_registration', 'Dice_after_registration', 'Jaccard_before', 'Jaccard_after',
          'Sensitivity_before', 'Sensitivity_after', 'Precision_before', 'Precision_after',
          'Specificity_before', 'Specificity_after', 'Accuracy_before', 'Accuracy_after']


def seg_summary(fx, moving, moved, sub_id):
    a, b = metrics.overlap_metrics(fx, moving), metrics.overlap_metrics(fx, moved)
    res = {'subject': sub_id}
    for key in ('dice', 'jaccard', 'sensitivity', 'precision', 'specificity', 'accuracy'):
        res['%s_before_registration' % key] = a[key]
        res['%s_after_registration' % key] = b[key]
    return res


def _load(path):
    path = path if len(path.split('.')) > 1 else path + '.nii.gz'
    return np.asarray(_nifti.load_nifti(path)[0], dtype=np.float64)


def main(argv=None):
    p = argparse.ArgumentParser(formatter_class=argparse.RawDescriptionHelpFormatter,
                                description='Evaluate the registration of two volumes')
    p.add_argument('--fx-seg-path', required=True, help='path to the spinal cord segmentation of the fixed image')
    p.add_argument('--moving-seg-path', required=True, help='path to the spinal cord segmentation of the moving image')
    p.add_argument('--warped-seg-path', required=True, help='path to the spinal cord segmentation of the moved image')
    p.add_argument('--sub-id', required=True, help='id of the subject')
    p.add_argument('--out-file', required=False, default='metrics_on_sc_seg.csv',
                   help='path to csv summarizing the results obtained on the SC segmentation with different metrics')
    p.add_argument('--append', type=int, required=False, default=1, choices=[0, 1],
                   help='Append results as a new line in the output csv file instead of overwriting it.')
    p.add_argument('--min-dice', required=False, type=int, default=0,
                   help='Minimum Dice score expected (percentage, to deal with int). If lower and not last-eval then '
                        'return a sys.exit(1) to signal this low score in the bash script and proceed to an '
                        "affine registration prior to the model's one")
    p.add_argument('--last-eval', type=int, required=False, default=1, choices=[0, 1],
                   help='Determine if this is the last evaluation that will be done (1) or not (0)')
    arg = p.parse_args(argv)

    res = seg_summary(_load(arg.fx_seg_path), _load(arg.moving_seg_path), _load(arg.warped_seg_path), arg.sub_id)
    if 100 * res['dice_after_registration'] < arg.min_dice and not arg.last_eval:
        return 1
    if not arg.append or not os.path.isfile(arg.out_file):
        with open(arg.out_file, 'w') as f:
            csv.DictWriter(f, fieldnames=HEADER).writeheader()
    with open(arg.out_file, 'a') as f:
        row = [datetime.datetime.now().strftime('%Y-%m-%d %H:%M:%S')] + [str(v) for v in res.values()]
        csv.writer(f, delimiter=',').writerow(row)
    return 0


if __name__ == '__main__':
    sys.exit(main())
