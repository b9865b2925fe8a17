"""Generate tests/golden/metrics_*.npz by EXECUTING the reference's own metric code.

Run in the build container only (needs /root/reference and scipy):
    python tests/golden/make_metrics_golden.py

* eval_reg_with_mi.py: the two functions `detect_zero_padding` and `normalized_mutual_information` are cut out of
  the file at run time (lines 16-74, everything between the imports and `if __name__`) and exec'ed with the
  names they need (np, scipy.stats.entropy); the script itself cannot be imported (nibabel at the top).
* eval_reg_on_sc_seg.py: lines 80-124 (TP .. Jaccard) are exec'ed with the three arrays bound.
Only seeded inputs and the outputs of the reference's arithmetic are stored -- no reference source.
"""
import os
import textwrap

import numpy as np
from scipy.stats import entropy

MI = '/root/reference/eval_reg_with_mi.py'
SC = '/root/reference/eval_reg_on_sc_seg.py'
HERE = os.path.dirname(os.path.abspath(__file__))


def mi_functions():
    with open(MI) as f:
        lines = f.readlines()
    src = ''.join(lines[15:74])
    assert src.lstrip().startswith('def detect_zero_padding') and 'def normalized_mutual_information' in src
    env = {'np': np, 'entropy': entropy}
    exec(compile(src, MI + ':16-74', 'exec'), env)
    return env['detect_zero_padding'], env['normalized_mutual_information']


def sc_metrics(fx, moving, moved):
    with open(SC) as f:
        lines = f.readlines()
    src = textwrap.dedent(''.join(lines[79:124]))
    assert src.lstrip().startswith('TP_moving') and 'jacc_fx_moved' in src
    src = src.replace('sys.exit(1)', 'pass')              # the early exit on a low Dice is script control flow
    env = {'np': np, 'fx_im_val': fx, 'moving_im_val': moving, 'moved_im_val': moved,
           'arg': type('A', (), {'min_dice': 0, 'last_eval': True})(), 'sys': None}
    exec(compile(src, SC + ':80-124', 'exec'), env)
    keys = ['TP', 'FP', 'TN', 'FN']
    out = {}
    for tag in ('moving', 'moved'):
        for k in keys:
            out['%s_%s' % (k, tag)] = np.float64(env['%s_%s' % (k, tag)])
        for k in ('dice', 'sens', 'spec', 'acc', 'prec', 'jacc'):
            out['%s_%s' % (k, tag)] = np.float64(env['%s_fx_%s' % (k, tag)])
    return out


def blobs(rng, shape, smooth):
    a = rng.standard_normal(shape)
    for _ in range(smooth):
        for ax in range(3):
            a = (np.roll(a, 1, ax) + a + np.roll(a, -1, ax)) / 3.0
    return a


def main():
    detect, nmi = mi_functions()
    cases = {'small': ((17, 13, 11), 2), 'mid': ((28, 24, 32), 3), 'flat_padded': ((24, 20, 16), 1)}
    for i, (name, (shape, smooth)) in enumerate(cases.items()):
        rng = np.random.default_rng(777 + i)
        fx = blobs(rng, shape, smooth)
        fx = (fx - fx.min()) / (fx.max() - fx.min())
        moving = np.clip(0.6 * fx + 0.4 * blobs(rng, shape, smooth) / 3 + 0.1, 0, None)
        moved = np.clip(0.9 * fx + 0.1 * blobs(rng, shape, smooth) / 3, 0, None)
        # zero padding as detect_zero_padding expects it (moving image zero outside a box)
        pad = np.zeros(shape, bool)
        pad[2:-3, 1:-2, 3:-1] = True
        moving = np.where(pad, moving + 0.01, 0.0)
        if name == 'flat_padded':
            moved = np.round(moved * 8) / 8                   # few distinct values: many samples on bin edges
            fx = fx.astype(np.float32).astype(np.float64)     # float32-representable inputs
            moving = moving.astype(np.float32).astype(np.float64)
            moved = moved.astype(np.float32).astype(np.float64)
        box = detect(moving)
        x0, y0, z0, x1, y1, z1 = box
        crop = lambda a: a[x0:x1 + 1, y0:y1 + 1, z0:z1 + 1]
        hist = np.histogramdd([crop(fx).ravel(), crop(moved).ravel()], bins=100)[0]
        res = dict(fx=fx, moving=moving, moved=moved, box=np.array(box, np.int64), hist_fx_moved=hist,
                   nmi_fx_moving=nmi(crop(fx), crop(moving)), nmi_fx_moved=nmi(crop(fx), crop(moved)),
                   nmi_moving_moved=nmi(crop(moving), crop(moved)), nmi_bins7=nmi(crop(fx), crop(moved), bins=7))
        # segmentations: thresholded blobs (values exactly 0 / 1, float64 like get_fdata())
        sfx = (fx > 0.55).astype(np.float64)
        smv = (np.roll(fx, 2, 0) > 0.55).astype(np.float64)
        smd = (np.roll(fx, 1, 1) > 0.5).astype(np.float64)
        res.update({'seg_fx': sfx, 'seg_moving': smv, 'seg_moved': smd})
        res.update({'sc_' + k: v for k, v in sc_metrics(sfx, smv, smd).items()})
        np.savez_compressed(os.path.join(HERE, 'metrics_%s.npz' % name), **res)
        print(name, shape, 'box', box, 'nmi', res['nmi_fx_moving'], res['nmi_fx_moved'], 'dice', res['sc_dice_moving'], res['sc_dice_moved'])


if __name__ == '__main__':
    main()
