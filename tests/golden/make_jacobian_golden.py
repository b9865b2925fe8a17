"""Generate tests/golden/jacobian_*.npz by EXECUTING the reference's own Jacobian lines.

Run in the build container only (needs /root/reference):
    python tests/golden/make_jacobian_golden.py

The reference script cannot be imported (it imports nibabel at the top and does all work
under ``__main__``), so this reads lines 62-78 of /root/reference/eval_reg_with_jacobian.py at
run time, dedents them and ``exec``s them with ``ddf`` bound to a seeded synthetic field.
No reference source is copied into this repository; only inputs (seed-derived, stored for
exactness) and the outputs of the reference's arithmetic are saved.
"""
import os
import textwrap

import numpy as np

REF = '/root/reference/eval_reg_with_jacobian.py'
HERE = os.path.dirname(os.path.abspath(__file__))


def run_reference_lines(ddf):
    with open(REF) as f:
        lines = f.readlines()
    src = textwrap.dedent(''.join(lines[61:78]))       # file lines 62..78 inclusive
    assert 'np.linalg.det' in src and 'height, width, depth' in src
    env = {'np': np, 'ddf': ddf}
    exec(compile(src, REF + ':62-78', 'exec'), env)
    return env['det'], env['negative_dets'], env['percentage_negative']


def smooth_field(rng, shape, std, smooth):
    """fp32 displacement field: white noise box-blurred `smooth` times, scaled to `std`."""
    f = rng.standard_normal(shape + (3,))
    for _ in range(smooth):
        for ax in range(3):
            f = (np.roll(f, 1, ax) + f + np.roll(f, -1, ax)) / 3.0
    f = f / f.std() * std
    return f.astype(np.float32)


def main():
    cases = {
        # name: (shape, std, smooth)
        'smooth_small': ((12, 11, 13), 0.4, 3),      # few folds
        'rough_folds': ((10, 12, 9), 1.5, 0),        # many folds (white noise)
        'min_size': ((5, 5, 5), 0.5, 0),             # a single interior voxel
        'anisotropic': ((24, 7, 16), 3.0, 1),
    }
    for i, (name, (shape, std, smooth)) in enumerate(cases.items()):
        rng = np.random.default_rng(20261018 + i)
        field = smooth_field(rng, shape, std, smooth)
        ddf = np.array(field[:, :, :, None, :], dtype=np.float64)   # get_fdata() -> float64
        det, n_neg, pct = run_reference_lines(ddf)
        np.savez_compressed(os.path.join(HERE, 'jacobian_%s.npz' % name),
                            field=field, det=det, n_neg=np.int64(n_neg), pct=np.float64(pct),
                            median=np.median(det), mean=np.mean(det), std=np.std(det))
        print(name, shape, 'n_neg', n_neg, 'of', det.size)


if __name__ == '__main__':
    main()
