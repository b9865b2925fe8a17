"""Device-level deformation ops on torch CUDA tensors, backed by libdfm.so through ctypes.

Tensors keep the reference's LOGICAL shape -- channels-last ``[B, X, Y, Z, C]`` -- while the
PHYSICAL layout is either contiguous channels-last ("cl") or channel-planar ("planar": the
storage is ``[B, C, X, Y, Z]`` and the tensor is a permuted view of it).  Kernels prefer
planar (128-bit coalesced access along z); inputs in either layout are consumed in place and
the layout bits are passed to the C ABI, so no transposition pass is needed at the boundary.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); all voxel work is
done by the CUDA kernels.  There is no CPU path: CPU tensors raise.
"""
import ctypes

import torch

from . import _coords, _lib

LINEAR, NEAREST = 'linear', 'nearest'
_INTERP = {LINEAR: _lib.DFM_LINEAR, NEAREST: _lib.DFM_NEAREST}


# ---------------------------------------------------------------------------------------
# layout helpers
# ---------------------------------------------------------------------------------------
def _interp_code(interp_method):
    try:
        return _INTERP[interp_method]
    except KeyError:
        raise ValueError("interp_method must be 'linear' or 'nearest', got %r" % (interp_method,))


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.DfmError('%s must be a CUDA tensor: the deformation engine has no CPU path' % name)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _perm_to_planar(nd):
    # logical [B, s1..sk, C] -> storage order [B, C, s1..sk]
    return (0, nd - 1) + tuple(range(1, nd - 1))


def _perm_to_logical(nd):
    return (0,) + tuple(range(2, nd)) + (1,)


def layout_of(t):
    """'cl', 'planar', 'both' (one channel, contiguous) or None (needs a copy)."""
    cl = t.is_contiguous()
    pl = t.permute(_perm_to_planar(t.dim())).is_contiguous()
    if cl and pl:
        return 'both'
    if cl:
        return 'cl'
    if pl:
        return 'planar'
    return None


def empty(shape, layout, device, dtype=torch.float32):
    """Uninitialised logical ``shape`` tensor with the given physical layout."""
    shape = tuple(int(s) for s in shape)
    if layout == 'cl':
        return torch.empty(shape, device=device, dtype=dtype)
    nd = len(shape)
    storage = torch.empty((shape[0], shape[-1]) + shape[1:-1], device=device, dtype=dtype)
    return storage.permute(_perm_to_logical(nd))


class _ToLayout(torch.autograd.Function):
    """A layout change is the identity on the logical tensor, so its gradient passes through."""

    @staticmethod
    def forward(ctx, t, layout):
        return _to_layout_raw(t, layout)

    @staticmethod
    def backward(ctx, g):
        return g, None


def to_layout(t, layout):
    """Return ``t`` (logical shape unchanged) in physical layout 'cl' or 'planar' using the
    library's own transposition kernels; no-op if it already is.  Differentiable."""
    cur = layout_of(t)
    if cur == 'both' or cur == layout:
        return t
    if torch.is_grad_enabled() and t.requires_grad:
        return _ToLayout.apply(t, layout)
    return _to_layout_raw(t, layout)


def _to_layout_raw(t, layout):
    t = t.detach()
    cur = layout_of(t)
    if cur == 'both' or cur == layout:
        return t
    if cur is None:
        t = t.contiguous()        # odd strides: normalise first (boundary glue, not hot path)
        if layout == 'cl':
            return t
    B, C = t.shape[0], t.shape[-1]
    N = 1
    for s in t.shape[1:-1]:
        N *= int(s)
    out = empty(t.shape, layout, t.device, t.dtype)
    fn = 'dfm_cl_to_planar' if layout == 'planar' else 'dfm_planar_to_cl'
    src, dst = (t, out)
    _lib.call(fn, _ptr(src), _ptr(dst), B, C, N, t.element_size(), _stream())
    return out


def _field_layout(t, name):
    """Ensure a 3-channel field is directly consumable; returns (tensor, is_cl)."""
    lay = layout_of(t)
    if lay is None:
        t = t.contiguous()
        lay = 'cl'
    return t, lay == 'cl'


def _check_field(t, name, nd=5):
    _require_cuda(t, name)
    if t.dim() != nd or t.shape[-1] != nd - 2:
        raise ValueError('%s must be [B, X, Y, Z, 3], got %s' % (name, tuple(t.shape)))
    if t.dtype != torch.float32:
        t = t.float()
    return t


# ---------------------------------------------------------------------------------------
# SpatialTransformer / transform
# ---------------------------------------------------------------------------------------
def _warp_fwd_raw(img, field, interp, fill_value, loc_absolute=False):
    B, Xi, Yi, Zi, C = img.shape
    _, X, Y, Z, _ = field.shape
    field, f_cl = _field_layout(field, 'field')
    lay = layout_of(img)
    if lay is None:
        img = img.contiguous()
        lay = 'cl'
    img_cl = lay == 'cl'
    out = empty((B, X, Y, Z, C), 'cl' if img_cl else 'planar', img.device, img.dtype)
    flags = ((_lib.FIELD_IN_CL if f_cl else 0) | (_lib.IMG_CL if img_cl else 0) |
             (_lib.LOC_ABSOLUTE if loc_absolute else 0))
    has_fill = fill_value is not None
    fill_f, fill_bits = 0.0, 0
    if has_fill:
        fill_f = float(fill_value)
        fill_bits = int.from_bytes(torch.tensor([fill_value]).to(img.dtype).numpy().tobytes(), 'little')
    _lib.call('dfm_warp_fwd', _ptr(img), _ptr(field), _ptr(out), B, C, Xi, Yi, Zi, X, Y, Z,
              _interp_code(interp), img.element_size(), int(has_fill), fill_f, fill_bits, flags, _stream())
    return out


class _WarpLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, field, fill_value):
        out = _warp_fwd_raw(img, field, LINEAR, fill_value)
        ctx.save_for_backward(img, field)
        ctx.has_fill = fill_value is not None
        return out

    @staticmethod
    def backward(ctx, gout):
        img, field = ctx.saved_tensors
        B, Xi, Yi, Zi, C = img.shape
        _, X, Y, Z, _ = field.shape
        field, f_cl = _field_layout(field, 'field')
        lay = layout_of(img)
        if lay is None:
            img = img.contiguous()
            lay = 'cl'
        img_cl = lay == 'cl'
        gout = to_layout(gout.float(), 'cl' if img_cl else 'planar')
        need_img, need_field = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gimg = gfield = None
        if need_img:
            gimg = empty(img.shape, 'cl' if img_cl else 'planar', img.device)
            gimg.zero_()
        if need_field:
            gfield = empty(field.shape, 'planar', img.device)
        flags = (_lib.FIELD_IN_CL if f_cl else 0) | (_lib.IMG_CL if img_cl else 0)
        _lib.call('dfm_warp_bwd', _ptr(gout), _ptr(img), _ptr(field), _ptr(gimg), _ptr(gfield),
                  B, C, Xi, Yi, Zi, X, Y, Z, int(ctx.has_fill), flags, _stream())
        return gimg, gfield, None


def warp(img, field, interp_method=LINEAR, fill_value=None, loc_absolute=False):
    """out[b, p, c] = interp(img[b, ..., c], p + field[b, p, :]).

    img [B, Xi, Yi, Zi, C]; field [B, X, Y, Z, 3] -> [B, X, Y, Z, C].  Mirrors the batched
    ``vxm.layers.SpatialTransformer`` of the reference (dfm.h: dfm_warp_fwd).
    loc_absolute: ``field`` holds sample locations themselves (``ne.utils.interpn``; no autograd)."""
    _require_cuda(img, 'img')
    field = _check_field(field, 'field')
    if img.dim() != 5:
        raise ValueError('img must be [B, X, Y, Z, C], got %s' % (tuple(img.shape),))
    if img.shape[0] != field.shape[0]:
        raise ValueError('batch mismatch: img %d vs field %d' % (img.shape[0], field.shape[0]))
    code = _interp_code(interp_method)
    if code == _lib.DFM_LINEAR:
        if img.dtype != torch.float32:
            img = img.float()
        # multi-channel images keep their layout: channels-last (the reference layout) runs the
        # lanes-over-channels kernel (dfm_warp_cl.cu), planar the TMA channel ring (dfm_brick_mc.cu)
        if not loc_absolute and torch.is_grad_enabled() and (img.requires_grad or field.requires_grad):
            return _WarpLinear.apply(img, field, fill_value)
        return _warp_fwd_raw(img.detach(), field.detach(), LINEAR, fill_value, loc_absolute)
    if img.element_size() not in (1, 2, 4, 8) or img.dtype == torch.bool:
        img = img.float()
    return _warp_fwd_raw(img.detach(), field.detach(), NEAREST, fill_value, loc_absolute)


class _WarpOneHot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, labels_u8, field, C, fill_value):
        B, Xi, Yi, Zi = labels_u8.shape
        _, X, Y, Z, _ = field.shape
        field, f_cl = _field_layout(field, 'field')
        out = torch.empty((B, X, Y, Z, C), device=field.device, dtype=torch.float32)
        has_fill = fill_value is not None
        _lib.call('dfm_warp_onehot_fwd', _ptr(labels_u8), _ptr(field), _ptr(out), B, C, Xi, Yi, Zi, X, Y, Z,
                  int(has_fill), float(fill_value or 0.0), _lib.FIELD_IN_CL if f_cl else 0, _stream())
        ctx.save_for_backward(labels_u8, field)
        ctx.C, ctx.has_fill, ctx.f_cl = C, has_fill, f_cl
        return out

    @staticmethod
    def backward(ctx, gout):
        labels_u8, field = ctx.saved_tensors
        if not ctx.needs_input_grad[1]:
            return None, None, None, None
        B, Xi, Yi, Zi = labels_u8.shape
        _, X, Y, Z, _ = field.shape
        gout = gout.float().contiguous()
        gfield = empty(field.shape, 'planar', field.device)
        _lib.call('dfm_warp_onehot_bwd', _ptr(gout), _ptr(labels_u8), _ptr(field), _ptr(gfield), B, ctx.C, Xi, Yi, Zi,
                  X, Y, Z, int(ctx.has_fill), _lib.FIELD_IN_CL if ctx.f_cl else 0, _stream())
        return None, gfield, None, None


ONEHOT_BWD_CMAX = 48


def warp_onehot(labels, field, num_labels, fill_value=None):
    """``warp(one_hot(labels, num_labels), field)`` -- the ``pred`` of train_synthmorph.py:298, whose first input is the one-hot
    map a ``labels_to_image`` generator emits -- computed from the LABEL MAP (dfm.h: dfm_warp_onehot_fwd / _bwd): a corner of a
    one-hot map is its label, so a voxel gathers 8 bytes instead of 8 x C floats and the one-hot tensor is never built.
    Bit-identical to the generic warp on the materialised tensor; differentiable with respect to ``field``.
    labels [B, X, Y, Z] or [B, X, Y, Z, 1], integer valued (any dtype); returns channels-last [B, X, Y, Z, num_labels]."""
    _require_cuda(labels, 'labels')
    field = _check_field(field, 'field')
    if labels.dim() == 5:
        if labels.shape[-1] != 1:
            raise ValueError('warp_onehot: labels must be [B, X, Y, Z] or [B, X, Y, Z, 1], got %s' % (tuple(labels.shape),))
        labels = labels[..., 0]
    if labels.dim() != 4 or labels.shape[0] != field.shape[0]:
        raise ValueError('warp_onehot: labels %s do not match field %s' % (tuple(labels.shape), tuple(field.shape)))
    C = int(num_labels)
    if not 1 <= C <= 256:
        raise ValueError('warp_onehot: num_labels must be in 1..256')
    lab = labels.detach()
    if lab.dtype != torch.uint8:
        lab = lab.clamp(0, 255).to(torch.uint8)       # values >= num_labels are all-zero rows, like a label outside the list
    lab = lab.contiguous()
    needs_grad = torch.is_grad_enabled() and field.requires_grad
    if needs_grad and C > ONEHOT_BWD_CMAX:            # the adjoint kernel stages a [256, C] gradient tile in shared memory
        onehot = torch.nn.functional.one_hot(lab.long(), 256)[..., :C].float()
        return warp(onehot, field, LINEAR, fill_value)
    if needs_grad:
        return _WarpOneHot.apply(lab, field, C, fill_value)
    return _WarpOneHot.forward(_NoCtx(), lab, field.detach(), C, fill_value)


class _NoCtx:
    """Stand-in for the autograd context when no gradient is needed."""

    def save_for_backward(self, *a):
        pass


def warp_channelwise(img, field, interp_method=LINEAR, fill_value=None, argmax=False):
    """Channel-wise transform: field [B, X, Y, Z, C, 3] carries one 3-vector per channel
    (``vxm.utils.transform`` as called at train_synthmorph.py:67).  Linear interpolation runs one kernel that
    addresses both tensors in the reference's channels-last layout (dfm.h: dfm_warp_channelwise_fwd);
    ``argmax=True`` also takes the ``tf.argmax(im, axis=-1)`` that follows at :68 inside the kernel and returns
    the uint8 label map [B, X, Y, Z].  Nearest interpolation (and volumes past the kernel's 32-bit index range)
    run as B*C one-channel warps."""
    _require_cuda(img, 'img')
    _require_cuda(field, 'field')
    if field.dim() != 6 or img.dim() != 5:
        raise ValueError('channel-wise transform: img [B, X, Y, Z, C], field [B, X, Y, Z, C, 3]')
    B, X, Y, Z, C, D = field.shape
    if D != 3 or img.shape[-1] != C or img.shape[0] != B:
        raise ValueError('channel-wise field must be [B, X, Y, Z, C, 3] with C == img channels')
    Xi, Yi, Zi = img.shape[1:4]
    if _interp_code(interp_method) == _lib.DFM_LINEAR:
        vol = img.detach().float().contiguous()                 # channels-last storage
        shift = field.detach().float().contiguous()
        out = torch.empty((B, X, Y, Z) if argmax else (B, X, Y, Z, C), device=img.device,
                          dtype=torch.uint8 if argmax else torch.float32)
        try:
            _lib.call('dfm_warp_channelwise_fwd', _ptr(vol), _ptr(shift), _ptr(out), B, C, Xi, Yi, Zi, X, Y, Z,
                      int(fill_value is not None), float(fill_value or 0.0), int(bool(argmax)), _stream())
            return out
        except _lib.DfmError as e:
            if e.code != _lib.DFM_EUNSUPPORTED:
                raise
    img_p = to_layout(img, 'planar')                         # storage [B, C, Xi, Yi, Zi]
    img_items = img_p.permute(0, 4, 1, 2, 3).reshape(B * C, Xi, Yi, Zi, 1)
    f_items = field.float().permute(0, 4, 5, 1, 2, 3).contiguous().reshape(B * C, 3, X, Y, Z).permute(0, 2, 3, 4, 1)   # planar items
    out = warp(img_items, f_items, interp_method, fill_value)               # [B*C, X, Y, Z, 1]
    out = out.reshape(B, C, X, Y, Z).permute(0, 2, 3, 4, 1)
    return out.argmax(-1).to(torch.uint8) if argmax else out


# ---------------------------------------------------------------------------------------
# field self/cross warps: SS step, compose, VecInt
# ---------------------------------------------------------------------------------------
def _field_warp_add_raw(src, own, scale, interp, out_layout='planar'):
    B, X, Y, Z, _ = own.shape
    _, Xs, Ys, Zs, _ = src.shape
    own, own_cl = _field_layout(own, 'own')
    if src is not own:
        src = to_layout(src, 'cl' if own_cl else 'planar')
    else:
        src = own
    out = empty(own.shape, out_layout, own.device)
    flags = (_lib.FIELD_IN_CL if own_cl else 0) | (_lib.FIELD_OUT_CL if out_layout == 'cl' else 0)
    _lib.call('dfm_field_warp_add', _ptr(src), _ptr(own), _ptr(out), B, Xs, Ys, Zs, X, Y, Z,
              float(scale), _interp_code(interp), flags, _stream())
    return out


class _Compose2(torch.autograd.Function):
    """out = own + interp(src, p + own)  (linear)."""

    @staticmethod
    def forward(ctx, src, own):
        ctx.save_for_backward(src, own)
        return _field_warp_add_raw(src, own, 1.0, LINEAR)

    @staticmethod
    def backward(ctx, gout):
        src, own = ctx.saved_tensors
        B, Xs, Ys, Zs, _ = src.shape
        _, X, Y, Z, _ = own.shape
        gout = to_layout(gout.float(), 'planar')
        src_p, own_p = to_layout(src, 'planar'), to_layout(own, 'planar')
        gsrc = empty(src.shape, 'planar', src.device)
        gsrc.zero_()
        gloc = empty(own.shape, 'planar', own.device)
        _lib.call('dfm_warp_bwd', _ptr(gout), _ptr(src_p), _ptr(own_p), _ptr(gsrc), _ptr(gloc),
                  B, 3, Xs, Ys, Zs, X, Y, Z, 0, 0, _stream())
        return gsrc, gout + gloc


def compose(transforms, interp_method=LINEAR, out_layout='planar'):
    """vxm.utils.compose for dense shifts: right fold ``curr = curr + transform(nxt, curr)``
    (bids_two_steps_registration.py:324,346,369,484).  Batched ``[B, X, Y, Z, 3]`` fields."""
    if len(transforms) < 2:
        raise ValueError('Compose transform list size must be greater than 1')
    curr = _check_field(transforms[-1], 'transforms[-1]')
    rest = list(reversed(transforms[:-1]))
    for n, nxt in enumerate(rest):
        nxt = _check_field(nxt, 'transforms[%d]' % (len(rest) - 1 - n))
        lay = out_layout if n == len(rest) - 1 else 'planar'
        if interp_method == LINEAR and torch.is_grad_enabled() and (nxt.requires_grad or curr.requires_grad):
            curr = _Compose2.apply(nxt, curr)
            if lay == 'cl':
                curr = to_layout(curr, 'cl')
        else:
            curr = _field_warp_add_raw(nxt, curr, 1.0, interp_method, lay)
    return curr


def _vecint_raw(svf, nsteps, save_steps, out_layout):
    B, X, Y, Z, _ = svf.shape
    svf, in_cl = _field_layout(svf, 'svf')
    out = empty(svf.shape, out_layout, svf.device)
    nbytes = _lib.load().dfm_vecint_workspace_bytes(B, X, Y, Z, nsteps, int(save_steps))
    work = torch.empty(max(nbytes // 4, 1), device=svf.device, dtype=torch.float32)
    flags = (_lib.FIELD_IN_CL if in_cl else 0) | (_lib.FIELD_OUT_CL if out_layout == 'cl' else 0)
    _lib.call('dfm_vecint_fwd', _ptr(svf), _ptr(out), _ptr(work), B, X, Y, Z, nsteps, int(save_steps),
              flags, _stream())
    return out, work


class _VecInt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, svf, nsteps):
        out, work = _vecint_raw(svf, nsteps, True, 'planar')
        ctx.save_for_backward(work)
        ctx.nsteps = nsteps
        ctx.shape = tuple(svf.shape)
        return out

    @staticmethod
    def backward(ctx, gout):
        (saved,) = ctx.saved_tensors
        B, X, Y, Z, _ = ctx.shape
        gout = to_layout(gout.float(), 'planar')
        if gout.data_ptr() % 16:
            gout = gout.clone(memory_format=torch.preserve_format)
        gsvf = empty(ctx.shape, 'planar', gout.device)
        n = B * 3 * X * Y * Z
        scratch = torch.empty(2 * n, device=gout.device, dtype=torch.float32)
        _lib.call('dfm_vecint_bwd', _ptr(gout), _ptr(saved), _ptr(gsvf), _ptr(scratch), B, X, Y, Z,
                  ctx.nsteps, _stream())
        return gsvf, None


def vecint(svf, nsteps=7, out_layout='planar'):
    """Scaling and squaring: ``v = svf / 2**n; n times v = v + interp(v, p + v)``.
    Mirrors vxm.layers.VecInt(method='ss', int_steps=nsteps) (dfm.h: dfm_vecint_fwd)."""
    svf = _check_field(svf, 'svf')
    nsteps = int(nsteps)
    if nsteps < 0:
        raise ValueError('nb_steps should be >= 0, found: %d' % nsteps)
    if torch.is_grad_enabled() and svf.requires_grad:
        out = _VecInt.apply(svf, nsteps)
        return to_layout(out, 'cl') if out_layout == 'cl' else out
    return _vecint_raw(svf, nsteps, False, out_layout)[0]


# ---------------------------------------------------------------------------------------
# resize / rescale
# ---------------------------------------------------------------------------------------
def _resize_raw(vol, out_shape, pre, post, interp, out_layout):
    B, Xi, Yi, Zi, C = vol.shape
    Xo, Yo, Zo = out_shape
    lay = layout_of(vol)
    if lay is None:
        vol = vol.contiguous()
        lay = 'cl'
    in_cl = lay == 'cl'
    dev = vol.device.index if vol.device.index is not None else torch.cuda.current_device()
    cx = _coords.device_tables(Xi, Xo, dev)[0]
    cy = _coords.device_tables(Yi, Yo, dev)[0]
    cz = _coords.device_tables(Zi, Zo, dev)[0]
    out = empty((B, Xo, Yo, Zo, C), out_layout, vol.device)
    flags = (_lib.FIELD_IN_CL if in_cl else 0) | (_lib.FIELD_OUT_CL if out_layout == 'cl' else 0)
    _lib.call('dfm_resize_fwd', _ptr(vol), _ptr(out), _ptr(cx), _ptr(cy), _ptr(cz), B, C, Xi, Yi, Zi,
              Xo, Yo, Zo, float(pre), float(post), _interp_code(interp), flags, _stream())
    return out


class _ResizeLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vol, out_shape, pre, post):
        ctx.in_shape = tuple(vol.shape)
        ctx.out_shape = tuple(out_shape)
        ctx.pre, ctx.post = pre, post
        return _resize_raw(vol, out_shape, pre, post, LINEAR, 'planar')

    @staticmethod
    def backward(ctx, gout):
        B, Xi, Yi, Zi, C = ctx.in_shape
        Xo, Yo, Zo = ctx.out_shape
        gout = to_layout(gout.float(), 'planar')
        dev = gout.device.index
        tx, ty, tz = (_coords.device_adjoint_taps(Xi, Xo, dev), _coords.device_adjoint_taps(Yi, Yo, dev),
                      _coords.device_adjoint_taps(Zi, Zo, dev))
        gin = empty(ctx.in_shape, 'planar', gout.device)
        work = None
        if max(tx[3], ty[3], tz[3]) >= 3:               # up-sampling adjoint: two separable passes through a workspace
            work = torch.empty(max(_lib.load().dfm_resize_bwd_workspace_bytes(B, C, Xi, Yi, Zo) // 4, 1), device=gout.device,
                               dtype=torch.float32)
        _lib.call('dfm_resize_bwd_ws', _ptr(gout), _ptr(gin), _ptr(work),
                  _ptr(tx[0]), _ptr(tx[1]), _ptr(tx[2]), tx[3], _ptr(ty[0]), _ptr(ty[1]), _ptr(ty[2]), ty[3],
                  _ptr(tz[0]), _ptr(tz[1]), _ptr(tz[2]), tz[3],
                  B, C, Xi, Yi, Zi, Xo, Yo, Zo, float(ctx.pre), float(ctx.post), _stream())
        return gin, None, None, None


def resize(vol, zoom_factor, interp_method=LINEAR, pre=1.0, post=1.0, out_layout='planar'):
    """ne.utils.resize on a batched ``[B, X, Y, Z, C]`` volume: corner-aligned resample onto
    ``linspace(0, n-1, int(n*zoom))`` per axis; out = post * interp(pre * vol)."""
    _require_cuda(vol, 'vol')
    if vol.dim() != 5:
        raise ValueError('vol must be [B, X, Y, Z, C], got %s' % (tuple(vol.shape),))
    if vol.dtype != torch.float32:
        vol = vol.float()
    zoom = list(zoom_factor) if isinstance(zoom_factor, (list, tuple)) else [zoom_factor] * 3
    out_shape = tuple(int(vol.shape[1 + d] * zoom[d]) for d in range(3))
    if _interp_code(interp_method) == _lib.DFM_LINEAR and torch.is_grad_enabled() and vol.requires_grad:
        out = _ResizeLinear.apply(vol, out_shape, float(pre), float(post))
        return to_layout(out, 'cl') if out_layout == 'cl' else out
    return _resize_raw(vol, out_shape, pre, post, interp_method, out_layout)


def rescale_dense_transform(trf, factor, interp_method=LINEAR, out_layout='planar'):
    """vxm.utils.rescale_dense_transform on batched fields: factor < 1 resizes then scales the
    vectors, factor >= 1 scales then resizes (3d_reg.py:394, bids_registration.py:398,
    bids_two_steps_registration.py:515)."""
    trf = _check_field(trf, 'transform')
    if factor < 1:
        return resize(trf, factor, interp_method, pre=1.0, post=factor, out_layout=out_layout)
    return resize(trf, factor, interp_method, pre=factor, post=1.0, out_layout=out_layout)


def rescale_warp(img, coarse_field, factor, fill_value=None, interp_method=LINEAR):
    """Fused ``RescaleTransform(factor)`` + ``SpatialTransformer`` of a one-channel image (dfm.h: dfm_rescale_warp_fwd,
    dfm_rescale_warp_nearest_fwd): the result of rescale_dense_transform followed by warp -- bit-identical wherever the
    marching kernels apply, within a few ulp of it in the default build otherwise -- without materialising the
    full-resolution field.  'nearest' moves 4-byte elements (float32 / int32 label maps).  Inference only (no autograd)."""
    _require_cuda(img, 'img')
    coarse_field = _check_field(coarse_field, 'coarse_field')
    if img.dim() != 5 or img.shape[-1] != 1:
        raise ValueError('rescale_warp: img must be [B, X, Y, Z, 1]')
    if factor < 1:
        raise ValueError('rescale_warp: factor must be >= 1')
    nearest = _interp_code(interp_method) == _lib.DFM_NEAREST
    if nearest:
        if img.element_size() != 4:
            img = img.float()
        img = img.contiguous()
    else:
        img = img.float().contiguous()
    coarse = to_layout(coarse_field.detach(), 'planar')
    B, Xi, Yi, Zi, _ = img.shape
    _, Xh, Yh, Zh, _ = coarse.shape
    X, Y, Z = int(Xh * factor), int(Yh * factor), int(Zh * factor)
    dev = img.device.index if img.device.index is not None else torch.cuda.current_device()
    cx = _coords.device_tables(Xh, X, dev)[0]
    cy = _coords.device_tables(Yh, Y, dev)[0]
    cz = _coords.device_tables(Zh, Z, dev)[0]
    out = torch.empty((B, X, Y, Z, 1), device=img.device, dtype=img.dtype)
    has_fill = fill_value is not None
    if nearest:
        fill_arg = int.from_bytes(torch.tensor([fill_value if has_fill else 0]).to(img.dtype).numpy().tobytes(), 'little')
        name = 'dfm_rescale_warp_nearest_fwd'
    else:
        fill_arg = float(fill_value or 0.0)
        name = 'dfm_rescale_warp_fwd'

    def call(work):
        _lib.call(name, _ptr(img), _ptr(coarse), _ptr(out), _ptr(cx), _ptr(cy), _ptr(cz), _ptr(work),
                  B, Xi, Yi, Zi, Xh, Yh, Zh, X, Y, Z, float(factor), int(has_fill), fill_arg, _stream())
    try:
        call(None)
    except _lib.DfmError as e:
        if e.code != _lib.DFM_EUNSUPPORTED:               # only "shape not covered by the fused kernel" falls back
            raise
        call(torch.empty(B * 3 * X * Y * Z, device=img.device, dtype=torch.float32))
    return out


# ---------------------------------------------------------------------------------------
# sub-volume stitching
# ---------------------------------------------------------------------------------------
def stitch_subvolumes(model_in_shape, im_shape, lst_coords_subvol, lst_warp_subvol, out_dtype=torch.float64):
    """``get_def_field_from_subvol`` of the reference scripts (3d_reg.py:214-259): pyramid-weighted
    average of overlapping tile fields.  ``lst_warp_subvol``: T arrays/tensors ``[tx, ty, tz, 3]`` (or one
    stacked ``[T, tx, ty, tz, 3]`` tensor); ``lst_coords_subvol``: T tuples (x_min, x_max, y_min, y_max,
    z_min, z_max).  Returns the channels-last field ``[X, Y, Z, 3]`` on the device (float64 like the
    reference by default)."""
    from . import _host
    tx, ty, tz = (int(d) for d in model_in_shape)
    X, Y, Z = (int(d) for d in im_shape[:3])
    if isinstance(lst_warp_subvol, torch.Tensor):
        tiles = _host.to_device(lst_warp_subvol, torch.float32)
    else:
        tiles = torch.stack([_host.to_device(w, torch.float32) for w in lst_warp_subvol], 0)
    T = tiles.shape[0]
    if len(lst_coords_subvol) != T or tuple(tiles.shape[1:]) != (tx, ty, tz, 3):
        raise ValueError('stitch_subvolumes: %d coordinate tuples for tiles of shape %s (expected [T, %d, %d, %d, 3])'
                         % (len(lst_coords_subvol), tuple(tiles.shape), tx, ty, tz))
    for c in lst_coords_subvol:
        if c[1] - c[0] != tx or c[3] - c[2] != ty or c[5] - c[4] != tz or c[0] < 0 or c[2] < 0 or c[4] < 0 \
                or c[1] > X or c[3] > Y or c[5] > Z:
            raise ValueError('stitch_subvolumes: tile %s does not match the tile shape or leaves the volume' % (c,))
    tiles, in_cl = _field_layout(tiles, 'tiles')
    mins = torch.tensor([[c[0], c[2], c[4]] for c in lst_coords_subvol], dtype=torch.int32).to(tiles.device)
    out = torch.empty((X, Y, Z, 3), device=tiles.device, dtype=out_dtype)
    flags = (_lib.FIELD_IN_CL if in_cl else 0) | _lib.FIELD_OUT_CL
    _lib.call('dfm_stitch_subvol', _ptr(tiles), _ptr(mins), _ptr(out), T, tx, ty, tz, X, Y, Z,
              int(out_dtype == torch.float64), flags, _stream())
    return out


# ---------------------------------------------------------------------------------------
# CUDA graphs
# ---------------------------------------------------------------------------------------
class Graphed:
    """Capture ``fn(*tensors)`` (a chain of libdfm launches on the current stream) in a CUDA graph and
    replay it: for single-volume calls the chain is launch-bound (9 kernels of ~10 us each), and a
    replay submits it with one driver call.  Inputs are copied into static buffers; the returned
    tensors are static too (valid until the next call) -- clone them to keep a result."""

    def __init__(self, fn, *example_inputs):
        cur = torch.cuda.current_stream()
        self.static_in = [t.detach().clone(memory_format=torch.preserve_format) for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():      # warm-up: lazy attribute / table set-up happens here
            for _ in range(2):
                fn(*self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for s, t in zip(self.static_in, inputs):
            if t.shape != s.shape:
                raise ValueError('graphed call needs the capture shapes %s, got %s' % (tuple(s.shape), tuple(t.shape)))
            s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_out


# ---------------------------------------------------------------------------------------
# Jacobian determinant
# ---------------------------------------------------------------------------------------
def jacobian_determinant(field, out_dtype=torch.float64, want_det=True, want_stats=True, exact=False):
    """det(I + grad u) on the interior [2:-2]^3 with 4th-order central differences, and the
    statistics of eval_reg_with_jacobian.py:62-91.

    field: [B, X, Y, Z, 3] (fp32 or fp64, either physical layout).
    Returns (det [B, X-4, Y-4, Z-4] or None, stats [B, 4] float64 = n_negative, sum, sum of
    squares, n_total; or None).
    Planar fp32 fields take the plane-marching kernel, whose stencils and 2x2 minors are fp32 (|error| < 1e-5): a
    determinant within that distance of zero can land on the other side of it than in the reference's all-float64
    evaluation.  ``exact=True`` evaluates in float64 like the reference whatever the layout (the fold count then
    matches ``np.linalg.det`` on the same values; ~8x slower)."""
    _require_cuda(field, 'field')
    if field.dim() != 5 or field.shape[-1] != 3:
        raise ValueError('field must be [B, X, Y, Z, 3], got %s' % (tuple(field.shape),))
    if field.dtype not in (torch.float32, torch.float64):
        field = field.float()
    B, X, Y, Z, _ = field.shape
    if exact and field.dtype == torch.float32:
        field = field.double()                              # the all-float64 kernel (eval_reg_with_jacobian.py:51: get_fdata())
    field, f_cl = _field_layout(field, 'field')
    det = torch.empty((B, X - 4, Y - 4, Z - 4), device=field.device, dtype=out_dtype) if want_det else None
    stats = part = None
    if want_stats:
        stats = torch.empty((B, 4), device=field.device, dtype=torch.float64)
        nbytes = _lib.load().dfm_jacdet_workspace_bytes(B, X, Y, Z)
        part = torch.empty(max(nbytes // 8, 1), device=field.device, dtype=torch.float64)
    _lib.call('dfm_jacdet', _ptr(field), _ptr(det), _ptr(stats), _ptr(part), B, X, Y, Z,
              int(field.dtype == torch.float64), int(out_dtype == torch.float64),
              _lib.FIELD_IN_CL if f_cl else 0, _stream())
    return det, stats


# --------------------------------------------------------------------------------------
# losses adjacent to the warp (dfm.h: dfm_dice_*, dfm_grad_l2_*)
# --------------------------------------------------------------------------------------
def _maps_layout(y_true, y_pred):
    """Bring two [B, ..., C] maps to ONE common physical layout; returns (true, pred, is_cl)."""
    lay = layout_of(y_pred)
    if lay is None:
        y_pred = y_pred.contiguous()
        lay = 'cl'
    if lay == 'both':
        lay = 'cl'
    return to_layout(y_true, lay), y_pred, lay == 'cl'


class _Dice(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_true, y_pred):
        y_true, y_pred, cl = _maps_layout(y_true.float(), y_pred.float())
        B, C = int(y_pred.shape[0]), int(y_pred.shape[-1])
        N = int(y_pred.numel() // max(B * C, 1))
        lib = _lib.load()
        sums = torch.empty((B, C, 2), device=y_pred.device, dtype=torch.float64)
        work = torch.empty(max(lib.dfm_dice_workspace_bytes(B, C, N) // 8, 1), device=y_pred.device, dtype=torch.float64)
        flags = _lib.IMG_CL if cl else 0
        _lib.call('dfm_dice_sums', _ptr(y_true), _ptr(y_pred), _ptr(sums), _ptr(work), B, C, N, flags, _stream())
        top, bottom = 2.0 * sums[..., 0], sums[..., 1]
        dice = torch.where(bottom != 0, top / torch.where(bottom != 0, bottom, torch.ones_like(bottom)), torch.zeros_like(top))
        ctx.save_for_backward(y_true, sums)
        ctx.cl, ctx.shape = cl, tuple(y_pred.shape)
        return (-dice.mean()).float()

    @staticmethod
    def backward(ctx, gout):
        y_true, sums = ctx.saved_tensors
        B, C = ctx.shape[0], ctx.shape[-1]
        N = int(y_true.numel() // max(B * C, 1))
        s0, s1 = sums[..., 0], sums[..., 1]
        ok = s1 != 0
        s1s = torch.where(ok, s1, torch.ones_like(s1))
        # loss = -(1/(B C)) sum 2 s0 / s1;  d s0/d p = t,  d s1/d p = 1
        k = -gout.double() / (B * C)
        alpha = torch.where(ok, k * 2.0 / s1s, torch.zeros_like(s1))
        beta = torch.where(ok, -k * 2.0 * s0 / (s1s * s1s), torch.zeros_like(s1))
        coef = torch.stack([alpha, beta], -1).float().contiguous()
        g = empty(ctx.shape, 'cl' if ctx.cl else 'planar', y_true.device)
        _lib.call('dfm_dice_bwd', _ptr(y_true), _ptr(coef), _ptr(g), B, C, N, _lib.IMG_CL if ctx.cl else 0, _stream())
        return None, g


def dice_loss(y_true, y_pred):
    """``vxm.losses.Dice().loss(y_true, y_pred)`` on [B, X, Y, Z, C] maps: -mean over (batch, channel) of
    divide_no_nan(2 sum(t p), sum(t + p)) (train_synthmorph.py:305).  Differentiable in y_pred."""
    _require_cuda(y_pred, 'y_pred')
    _require_cuda(y_true, 'y_true')
    if tuple(y_true.shape) != tuple(y_pred.shape):
        raise ValueError('dice_loss: shapes differ: %s vs %s' % (tuple(y_true.shape), tuple(y_pred.shape)))
    return _Dice.apply(y_true, y_pred)


class _GradL2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, loss_mult):
        flow, cl = _field_layout(flow.float(), 'flow')
        B, X, Y, Z, _ = flow.shape
        lib = _lib.load()
        sums = torch.empty((B, 3), device=flow.device, dtype=torch.float64)
        work = torch.empty(max(lib.dfm_grad_l2_workspace_bytes(B, X, Y, Z) // 8, 1), device=flow.device, dtype=torch.float64)
        _lib.call('dfm_grad_l2_sums', _ptr(flow), _ptr(sums), _ptr(work), B, X, Y, Z, _lib.FIELD_IN_CL if cl else 0, _stream())
        counts = torch.tensor([(X - 1) * Y * Z * 3, X * (Y - 1) * Z * 3, X * Y * (Z - 1) * 3], device=flow.device, dtype=torch.float64)
        ctx.save_for_backward(flow, counts)
        ctx.cl, ctx.mult = cl, float(loss_mult)
        return ((sums / counts).sum(-1) / 3.0 * float(loss_mult)).float()       # per batch item, like the reference

    @staticmethod
    def backward(ctx, gout):
        flow, counts = ctx.saved_tensors
        B, X, Y, Z, _ = flow.shape
        coef = (2.0 * ctx.mult / 3.0 * gout.double().reshape(B, 1) / counts.reshape(1, 3)).float().contiguous()
        g = empty(flow.shape, 'cl' if ctx.cl else 'planar', flow.device)
        _lib.call('dfm_grad_l2_bwd', _ptr(flow), _ptr(coef), _ptr(g), B, X, Y, Z, _lib.FIELD_IN_CL if ctx.cl else 0, _stream())
        return g, None


def grad_l2_loss(flow, loss_mult=1.0):
    """``vxm.losses.Grad('l2', loss_mult).loss(None, flow)`` on a [B, X, Y, Z, 3] field: per batch item the
    mean over the three axes of the mean squared forward difference (train_synthmorph.py:306)."""
    flow = _check_field(flow, 'flow')
    if min(flow.shape[1:4]) < 2:
        raise ValueError('grad_l2_loss: every spatial axis needs at least 2 voxels')
    return _GradL2.apply(flow, loss_mult)


class _WarpDice(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, field, y_true, fill_value):
        B, Xi, Yi, Zi, C = img.shape
        _, X, Y, Z, _ = field.shape
        field, f_cl = _field_layout(field, 'field')
        # forward: the stand-alone warp and a streaming Dice pass (fusing them was measured and is slower);
        # the warped map is a temporary, only the B*C*2 sums are kept for the backward pass
        pred = _warp_fwd_raw(img, field, LINEAR, fill_value)
        lib = _lib.load()
        sums = torch.empty((B, C, 2), device=img.device, dtype=torch.float64)
        N = X * Y * Z
        work = torch.empty(max(lib.dfm_dice_workspace_bytes(B, C, N) // 8, 1), device=img.device, dtype=torch.float64)
        _lib.call('dfm_dice_sums', _ptr(y_true), _ptr(pred), _ptr(sums), _ptr(work), B, C, N, _lib.IMG_CL, _stream())
        del pred
        top, bottom = 2.0 * sums[..., 0], sums[..., 1]
        dice = torch.where(bottom != 0, top / torch.where(bottom != 0, bottom, torch.ones_like(bottom)), torch.zeros_like(top))
        ctx.save_for_backward(img, field, y_true, sums)
        ctx.f_cl, ctx.has_fill = f_cl, fill_value is not None
        return (-dice.mean()).float()

    @staticmethod
    def backward(ctx, gout):
        img, field, y_true, sums = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise NotImplementedError('warp_dice_loss: gradient with respect to the moving map is not fused; use warp + dice_loss')
        B, Xi, Yi, Zi, C = img.shape
        _, X, Y, Z, _ = field.shape
        s0, s1 = sums[..., 0], sums[..., 1]
        ok = s1 != 0
        s1s = torch.where(ok, s1, torch.ones_like(s1))
        k = -gout.double() / (B * C)
        coef = torch.stack([torch.where(ok, k * 2.0 / s1s, torch.zeros_like(s1)),
                            torch.where(ok, -k * 2.0 * s0 / (s1s * s1s), torch.zeros_like(s1))], -1).float().contiguous()
        gfield = empty(field.shape, 'planar', img.device)
        _lib.call('dfm_warp_dice_bwd', _ptr(y_true), _ptr(coef), _ptr(img), _ptr(field), _ptr(gfield), B, C, Xi, Yi, Zi,
                  X, Y, Z, int(ctx.has_fill), _lib.FIELD_IN_CL if ctx.f_cl else 0, _stream())
        return None, gfield, None, None


def warp_dice_loss(img, field, y_true, fill_value=None):
    """``Dice().loss(y_true, SpatialTransformer('linear')([img, field]))`` on channels-last maps
    (train_synthmorph.py:298 + :305) without keeping the warped map or materialising the Dice gradient: the
    backward pass is ONE kernel that forms d Dice / d pred on the fly (dfm.h: dfm_warp_dice_bwd).
    Differentiable in ``field``.  Channel counts the fused kernel does not cover (odd C > 32, C > 64, planar
    maps) run the two stand-alone ops."""
    _require_cuda(img, 'img')
    _require_cuda(y_true, 'y_true')
    field = _check_field(field, 'field')
    C = int(img.shape[-1])
    fused = (img.dim() == 5 and 2 <= C <= 64 and (C <= 32 or C % 2 == 0) and layout_of(img) in ('cl', 'both') and
             layout_of(y_true) in ('cl', 'both') and img.dtype == torch.float32 and y_true.dtype == torch.float32 and
             min(img.shape[1:4]) >= 2 and not img.requires_grad and
             tuple(y_true.shape) == (img.shape[0],) + tuple(field.shape[1:4]) + (C,))
    if not fused:
        return dice_loss(y_true, warp(img, field, LINEAR, fill_value))
    return _WarpDice.apply(img, field, y_true, fill_value)
