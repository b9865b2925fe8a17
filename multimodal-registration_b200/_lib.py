"""ctypes binding of libdfm.so (the C ABI declared in include/dfm.h).

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# two builds of the same sources (csrc/Makefile): fused/packed accumulation (default) and the
# reference's op order with separately rounded ops (bit-identical to the oracle)
LIB_PATHS = {False: os.path.join(_HERE, 'libdfm.so'), True: os.path.join(_HERE, 'libdfm_exact.so')}
LIB_PATH = LIB_PATHS[False]

DFM_LINEAR, DFM_NEAREST = 0, 1
FIELD_IN_CL, FIELD_OUT_CL, IMG_CL, LOC_ABSOLUTE = 1, 2, 4, 8

_c = ctypes
_p, _i, _f, _u, _z, _u64 = _c.c_void_p, _c.c_int, _c.c_float, _c.c_uint, _c.c_size_t, _c.c_uint64

# name -> (restype, argtypes); must list every symbol include/dfm.h declares
SIGNATURES = {
    'dfm_version': (_i, []),
    'dfm_exact_order': (_i, []),
    'dfm_last_error': (_c.c_char_p, []),
    'dfm_warp_fwd': (_i, [_p, _p, _p] + [_i] * 8 + [_i, _i, _i, _f, _u64, _u, _p]),
    'dfm_warp_channelwise_fwd': (_i, [_p, _p, _p] + [_i] * 9 + [_f, _i, _p]),
    'dfm_rescale_warp_fwd': (_i, [_p] * 7 + [_i] * 10 + [_f, _i, _f, _p]),
    'dfm_rescale_warp_nearest_fwd': (_i, [_p] * 7 + [_i] * 10 + [_f, _i, _c.c_uint32, _p]),
    'dfm_warp_onehot_fwd': (_i, [_p, _p, _p] + [_i] * 8 + [_i, _f, _u, _p]),
    'dfm_warp_onehot_bwd': (_i, [_p] * 4 + [_i] * 8 + [_i, _u, _p]),
    'dfm_warp_bwd': (_i, [_p] * 5 + [_i] * 8 + [_i, _u, _p]),
    'dfm_field_warp_add': (_i, [_p, _p, _p] + [_i] * 7 + [_f, _i, _u, _p]),
    'dfm_vecint_workspace_bytes': (_z, [_i] * 6),
    'dfm_vecint_fwd': (_i, [_p, _p, _p] + [_i] * 6 + [_u, _p]),
    'dfm_vecint_bwd': (_i, [_p] * 4 + [_i] * 5 + [_p]),
    'dfm_ss_step_bwd': (_i, [_p] * 3 + [_i] * 4 + [_f, _p]),
    'dfm_ss_step_bwd_bounded': (_i, [_p] * 4 + [_f] + [_i] * 4 + [_f, _p]),
    'dfm_resize_fwd': (_i, [_p] * 5 + [_i] * 8 + [_f, _f, _i, _u, _p]),
    'dfm_resize_bwd': (_i, [_p, _p] + [_p, _p, _p, _i] * 3 + [_i] * 8 + [_f, _f, _p]),
    'dfm_resize_bwd_workspace_bytes': (_z, [_i] * 5),
    'dfm_resize_bwd_ws': (_i, [_p, _p, _p] + [_p, _p, _p, _i] * 3 + [_i] * 8 + [_f, _f, _p]),
    'dfm_jacdet_workspace_bytes': (_z, [_i] * 4),
    'dfm_jacdet': (_i, [_p] * 4 + [_i] * 6 + [_u, _p]),
    'dfm_stitch_subvol': (_i, [_p, _p, _p] + [_i] * 8 + [_u, _p]),
    'dfm_dice_workspace_bytes': (_z, [_i, _i, _z]),
    'dfm_dice_sums': (_i, [_p, _p, _p, _p, _i, _i, _z, _u, _p]),
    'dfm_dice_bwd': (_i, [_p, _p, _p, _i, _i, _z, _u, _p]),
    'dfm_warp_dice_bwd': (_i, [_p] * 5 + [_i] * 8 + [_i, _u, _p]),
    'dfm_grad_l2_workspace_bytes': (_z, [_i] * 4),
    'dfm_grad_l2_sums': (_i, [_p, _p, _p] + [_i] * 4 + [_u, _p]),
    'dfm_grad_l2_bwd': (_i, [_p, _p, _p] + [_i] * 4 + [_u, _p]),
    'dfm_metrics_workspace_bytes': (_z, []),
    'dfm_minmax': (_i, [_p, _z, _i, _p, _p, _p]),
    'dfm_joint_hist': (_i, [_p, _p, _z, _i, _p, _p, _i, _p, _p]),
    'dfm_axis_sums': (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p]),
    'dfm_overlap_sums': (_i, [_p, _p, _z, _i, _p, _p, _p]),
    'dfm_synth_intensity': (_i, [_p, _p, _p, _i, _u64, _p, _z, _p]),
    'dfm_conv1d_axis': (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _i, _p]),
    'dfm_scale_exp_clip': (_i, [_p, _p, _p, _z, _f, _f, _p]),
    'dfm_norm_gamma': (_i, [_p, _p, _p, _p, _i, _z, _p]),
    'dfm_onehot': (_i, [_p, _p, _i, _i, _p, _z, _p]),
    'dfm_cl_to_planar': (_i, [_p, _p, _i, _i, _z, _i, _p]),
    'dfm_planar_to_cl': (_i, [_p, _p, _i, _i, _z, _i, _p]),
}


DFM_OK, DFM_EINVAL, DFM_EALIGN, DFM_EUNSUPPORTED, DFM_ECUDA = 0, -1, -2, -3, -4


class DfmError(RuntimeError):
    """Raised for every failed library call; ``code`` is the DFM_E* return value (None for loader errors)."""

    def __init__(self, msg, code=None):
        super().__init__(msg)
        self.code = code


_libs = {}
_exact = os.environ.get('DFM_EXACT', '0') not in ('', '0')


def use(exact):
    """Select the arithmetic mode for subsequent calls: False = libdfm.so (fused/packed, default),
    True = libdfm_exact.so (reference op order, bit-identical to the oracle).  The environment
    variable DFM_EXACT=1 selects the exact build at import time."""
    global _exact
    _exact = bool(exact)
    return load()


def exact_order():
    return bool(load().dfm_exact_order())


def load():
    """Load the selected library once; raise loudly if it has not been built."""
    lib = _libs.get(_exact)
    if lib is not None:
        return lib
    path = LIB_PATHS[_exact]
    if not os.path.exists(path):
        raise DfmError(
            '%s not found -- build it with `python __graft_entry__.py` or '
            '`make -C multimodal-registration_b200/csrc`. There is no CPU fallback.' % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if bool(lib.dfm_exact_order()) != _exact:
        raise DfmError('%s reports the wrong arithmetic mode' % path)
    _libs[_exact] = lib
    return lib


def call(name, *args):
    """Call an int-returning entry point and raise DfmError with dfm_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise DfmError('%s failed (%d): %s' % (name, rc, lib.dfm_last_error().decode()), rc)


def version():
    return load().dfm_version()
