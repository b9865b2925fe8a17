"""CPU-side checks of the boundary: libdfm.so loads, exports every symbol include/dfm.h declares,
validates arguments without touching a GPU, and the host shim refuses to run without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import _coords, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'dfm.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dfm_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), 'libdfm.so does not export %s' % n
    assert sorted(_lib.SIGNATURES) == names, 'ctypes signature table out of sync with dfm.h'
    assert lib.dfm_version() == 100


def test_argument_validation_without_gpu():
    lib = _lib.load()
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)
    # bad interpolation code
    rc = lib.dfm_field_warp_add(one, one, ctypes.c_void_p(32), 1, 4, 4, 4, 4, 4, 4, 1.0, 7, 0, null)
    assert rc == -1 and b'interp' in lib.dfm_last_error()
    # non power-of-two scale
    rc = lib.dfm_field_warp_add(one, one, ctypes.c_void_p(32), 1, 4, 4, 4, 4, 4, 4, 0.3, 0, 0, null)
    assert rc == -1 and b'power of two' in lib.dfm_last_error()
    # jacobian needs >= 5 voxels per axis (interior [2:-2])
    rc = lib.dfm_jacdet(one, one, null, null, 1, 4, 8, 8, 0, 0, 0, null)
    assert rc == -1 and b'at least 5' in lib.dfm_last_error()
    # linear warp of a non-fp32 image
    rc = lib.dfm_warp_fwd(one, one, ctypes.c_void_p(32), 1, 1, 4, 4, 4, 4, 4, 4, 0, 1, 0, 0.0, 0, 0, null)
    assert rc == -1 and b'fp32' in lib.dfm_last_error()
    # empty batch is a no-op
    assert lib.dfm_vecint_fwd(one, ctypes.c_void_p(32), null, 0, 4, 4, 4, 7, 0, 0, null) == 0
    assert lib.dfm_vecint_workspace_bytes(2, 4, 4, 4, 7, 0) == 2 * 3 * 64 * 4 + 768     # ping-pong field + three per-item maxima
    assert lib.dfm_vecint_workspace_bytes(2, 4, 4, 4, 7, 1) == 7 * 2 * 3 * 64 * 4 + 768
    assert lib.dfm_vecint_workspace_bytes(2, 4, 4, 4, 1, 0) == 0


def test_new_entry_points_validate_without_gpu():
    """dfm_rescale_warp_nearest_fwd / dfm_warp_onehot_fwd / _bwd reject bad arguments before touching the device, and the
    Python wrappers refuse CPU tensors (no CPU path)."""
    lib = _lib.load()
    null, one, two = ctypes.c_void_p(0), ctypes.c_void_p(512), ctypes.c_void_p(1024)
    # factor < 1: the fused call only covers the scale-then-resize order
    rc = lib.dfm_rescale_warp_nearest_fwd(one, one, two, one, one, one, null, 1, 8, 8, 8, 4, 4, 4, 8, 8, 8, 0.5, 0, 0, null)
    assert rc == -1 and b'factor' in lib.dfm_last_error()
    # out aliasing img
    rc = lib.dfm_rescale_warp_nearest_fwd(one, one, one, one, one, one, null, 1, 8, 8, 8, 4, 4, 4, 8, 8, 8, 2.0, 0, 0, null)
    assert rc == -1 and b'alias' in lib.dfm_last_error()
    # empty batch is a no-op
    assert lib.dfm_rescale_warp_nearest_fwd(one, one, two, one, one, one, null, 0, 8, 8, 8, 4, 4, 4, 8, 8, 8, 2.0, 0, 0, null) == 0
    # one-hot warp: labels are bytes, the label volume needs every axis >= 2, the adjoint's gradient tile bounds C
    rc = lib.dfm_warp_onehot_fwd(one, one, two, 1, 300, 8, 8, 8, 8, 8, 8, 0, 0.0, 0, null)
    assert rc == -1 and b'byte' in lib.dfm_last_error()
    rc = lib.dfm_warp_onehot_fwd(one, one, two, 1, 26, 1, 8, 8, 8, 8, 8, 0, 0.0, 0, null)
    assert rc == -1 and b'axis' in lib.dfm_last_error()
    rc = lib.dfm_warp_onehot_bwd(one, one, one, two, 1, 60, 8, 8, 8, 8, 8, 8, 0, 0, null)
    assert rc == -3 and b'tile' in lib.dfm_last_error()
    assert lib.dfm_warp_onehot_fwd(one, one, two, 0, 26, 8, 8, 8, 8, 8, 8, 0, 0.0, 0, null) == 0
    from multimodal_registration_b200 import ops
    lab = torch.zeros((1, 4, 4, 4), dtype=torch.uint8)
    fld = torch.zeros((1, 4, 4, 4, 3))
    with pytest.raises(_lib.DfmError):
        ops.warp_onehot(lab, fld, 5)
    with pytest.raises(_lib.DfmError):
        ops.rescale_warp(torch.zeros((1, 8, 8, 8, 1)), fld, 2, None, 'nearest')


def test_coordinate_tables_follow_tf_linspace():
    c = _coords.linspace_tf(80, 160)
    assert c.dtype == np.float32 and c[0] == 0 and c[-1] == 79
    d = np.float32(np.float32(79) / np.float32(159))
    np.testing.assert_array_equal(c[1:-1], (d * np.arange(1, 159).astype(np.float32)).astype(np.float32))
    np.testing.assert_array_equal(_coords.linspace_tf(7, 7), np.arange(7, dtype=np.float32))
    from oracle import interp_oracle as io
    for n_in, n_out in [(80, 160), (96, 192), (160, 80), (5, 13), (9, 1), (3, 3)]:
        np.testing.assert_array_equal(_coords.linspace_tf(n_in, n_out), io.linspace_tf(0., n_in - 1., n_out))


def test_support_ranges_cover_exactly_the_taps():
    for n_in, n_out in [(8, 16), (16, 8), (5, 13), (6, 6)]:
        c = _coords.linspace_tf(n_in, n_out)
        lo, hi = _coords.support_ranges(c, n_in)
        i0 = np.clip(np.floor(c), 0, n_in - 1).astype(int)
        i1 = np.minimum(i0 + 1, n_in - 1)
        for i in range(n_in):
            taps = [j for j in range(n_out) if i0[j] == i or i1[j] == i]
            if taps:
                assert lo[i] == taps[0] and hi[i] == taps[-1] + 1
                assert taps == list(range(lo[i], hi[i]))      # contiguous
            else:
                assert lo[i] == hi[i]


def test_adjoint_taps_are_the_transpose_of_the_forward_weights():
    for n_in, n_out in [(8, 16), (80, 160), (16, 8), (5, 13), (6, 6), (1, 4), (4, 1)]:
        c = _coords.linspace_tf(n_in, n_out)
        lo, cnt, w = _coords.adjoint_taps(c, n_in)
        A = np.zeros((n_out, n_in))
        cl = np.clip(c, 0, n_in - 1)
        i1 = np.minimum(np.floor(cl).astype(int) + 1, n_in - 1)
        i0 = np.maximum(i1 - 1, 0)
        w0 = (i1.astype(np.float32) - cl).astype(np.float32)
        for j in range(n_out):
            A[j, i0[j]] += w0[j]
            A[j, i1[j]] += np.float32(1) - w0[j]
        Bm = np.zeros_like(A)
        for i in range(n_in):
            for k in range(cnt[i]):
                Bm[lo[i] + k, i] = w[i, k]
        np.testing.assert_allclose(A, Bm, atol=1e-7)
        np.testing.assert_allclose(A.sum(1), 1.0, atol=1e-6)        # every output is a convex combination


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_no_cpu_fallback():
    vxm = mrb.voxelmorph
    with pytest.raises(_lib.DfmError):
        vxm.utils.transform(np.zeros((4, 4, 4, 1), np.float32), np.zeros((4, 4, 4, 3), np.float32))
    with pytest.raises(_lib.DfmError):
        mrb.ops.vecint(torch.zeros(1, 4, 4, 4, 3))


def test_shim_signatures_and_errors():
    vxm, ne = mrb.voxelmorph, mrb.neurite
    import inspect
    sig = inspect.signature(vxm.layers.SpatialTransformer.__init__)
    assert list(sig.parameters)[1:6] == ['interp_method', 'indexing', 'single_transform', 'fill_value', 'shift_center']
    assert inspect.signature(vxm.layers.VecInt.__init__).parameters['int_steps'].default == 7
    assert list(inspect.signature(vxm.utils.transform).parameters) == ['vol', 'loc_shift', 'interp_method', 'indexing', 'fill_value']
    assert list(inspect.signature(vxm.utils.compose).parameters) == ['transforms', 'interp_method', 'shift_center', 'indexing']
    assert list(inspect.signature(vxm.networks.Transform.__init__).parameters)[1:] == ['inshape', 'affine', 'interp_method', 'rescale', 'fill_value', 'nb_feats']
    assert list(inspect.signature(ne.utils.interpn).parameters) == ['vol', 'loc', 'interp_method', 'fill_value']
    with pytest.raises(ValueError):
        vxm.utils.compose([np.zeros((4, 4, 4, 3))])
    with pytest.raises(ValueError):
        vxm.utils.compose([np.zeros((4, 4, 4, 3))] * 2, indexing='xy')
    with pytest.raises(NotImplementedError):
        vxm.layers.VecInt(method='ode')
    t = vxm.networks.Transform((8, 8, 8), rescale=2)
    assert t.trf_shape == (4, 4, 4)
    m = vxm.networks.VxmDense((160, 160, 192), int_steps=5, svf_resolution=2, int_resolution=2)
    assert m.svf_size == (80, 80, 96) and m.int_size == (80, 80, 96)
