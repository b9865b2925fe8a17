#!/usr/bin/env python
"""Drop-in for the reference's eval_reg_with_mi.py (same flags, same CSV columns): normalised mutual
information between the fixed, moving and moved images, with the voxel passes on the GPU
(multimodal_registration_b200.metrics: dfm_minmax, dfm_joint_hist, dfm_axis_sums).

Reference behaviour kept (eval_reg_with_mi.py:100-160): volumes are read as float64 (get_fdata()), cropped to
the non-zero-padded box of the MOVING image, NMI with 100 bins per axis, percentage improvement rounded to two
decimals, one CSV row per call, header written when the file is new or --append 0.
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

HEADER = ['Timestamp', 'Subject', 'NMI_before_registration', 'NMI_after_registration',
          'NMI_between_moving_and_moved_images', 'Percentage_nmi_improvement_registration']


def nmi_summary(fx, moving, moved, sub_id):
    x_min, y_min, z_min, x_max, y_max, z_max = metrics.detect_zero_padding(moving)
    fx, moving, moved = (np.ascontiguousarray(a[x_min:x_max + 1, y_min:y_max + 1, z_min:z_max + 1]) for a in (fx, moving, moved))
    nmi_fx_moving = metrics.normalized_mutual_information(fx, moving)
    nmi_fx_moved = metrics.normalized_mutual_information(fx, moved)
    nmi_moving_moved = metrics.normalized_mutual_information(moving, moved)
    perc = 100 * (nmi_fx_moved - nmi_fx_moving) / nmi_fx_moving
    return {'subject': sub_id, 'nmi_before_registration': nmi_fx_moving, 'nmi_after_registration': nmi_fx_moved,
            'nmi_between_moving_and_moved_images': nmi_moving_moved,
            'perc_nmi_improvement_with_registration': np.round(perc, 2)}


def _load(path):
    path = path if len(path.split('.')) > 1 else path + '.nii.gz'
    return np.asarray(_nifti.load_nifti(path)[0], dtype=np.float64)      # get_fdata()


def main(argv=None):
    p = argparse.ArgumentParser(formatter_class=argparse.RawDescriptionHelpFormatter,
                                description='Evaluate the registration of two volumes')
    p.add_argument('--fx-im-path', required=True, help='path to the fixed image')
    p.add_argument('--moving-im-path', required=True, help='path to the moving image')
    p.add_argument('--warped-im-path', required=True, help='path to the moved image')
    p.add_argument('--sub-id', required=True, help='id of the subject')
    p.add_argument('--out-file', required=False, default='nmi.csv',
                   help='path to csv summarizing the mutual information results')
    p.add_argument('--append', type=int, required=False, default=1, choices=[0, 1],
                   help='Append results as a new line in the output csv file instead of overwriting it.')
    arg = p.parse_args(argv)

    res = nmi_summary(_load(arg.fx_im_path), _load(arg.moving_im_path), _load(arg.warped_im_path), arg.sub_id)
    if not arg.append or not os.path.isfile(arg.out_file):
        with open(arg.out_file, 'w') as f:
            csv.DictWriter(f, fieldnames=HEADER).writeheader()
    with open(arg.out_file, 'a') as f:
        row = [datetime.datetime.now().strftime('%Y-%m-%d %H:%M:%S')] + [str(v) for v in res.values()]
        csv.writer(f, delimiter=',').writerow(row)
    return 0


if __name__ == '__main__':
    sys.exit(main())
