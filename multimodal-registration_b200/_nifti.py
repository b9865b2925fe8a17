"""Minimal NIfTI-1 reader / writer (.nii, .nii.gz) -- nibabel is not available in this image.

Covers what the hot-path scripts need: single-file NIfTI-1, little or
big endian, the common scalar dtypes, up to 7 dimensions, sform/qform affine, `intent_code`
(1007 = vector, used for SCT warps: 3d_reg.py:419, bids_registration.py:423).  Data is returned in
Fortran order reshaped to the NIfTI `dim` (like nibabel's `get_fdata` / `dataobj`), with
scl_slope / scl_inter applied when set.
"""
import gzip
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8,
           512: np.uint16, 768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).name: k for k, v in _DTYPES.items()}


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith('.gz') else open(path, mode)


def _quat_affine(b, c, d, qx, qy, qz, dx, dy, dz, qfac):
    a = np.sqrt(max(1.0 - (b * b + c * c + d * d), 0.0))
    R = np.array([[a * a + b * b - c * c - d * d, 2 * b * c - 2 * a * d, 2 * b * d + 2 * a * c],
                  [2 * b * c + 2 * a * d, a * a + c * c - b * b - d * d, 2 * c * d - 2 * a * b],
                  [2 * b * d - 2 * a * c, 2 * c * d + 2 * a * b, a * a + d * d - c * c - b * b]])
    aff = np.eye(4)
    aff[:3, :3] = R * np.array([dx, dy, dz * (qfac if qfac else 1.0)])
    aff[:3, 3] = [qx, qy, qz]
    return aff


def load_nifti(path, return_header=False):
    """Returns (data, affine[, header dict])."""
    with _open(path, 'rb') as f:
        raw = f.read()
    if len(raw) < 348:
        raise ValueError('%s: not a NIfTI-1 file (too short)' % path)
    endian = '<' if struct.unpack('<i', raw[:4])[0] == 348 else '>'
    if struct.unpack(endian + 'i', raw[:4])[0] != 348:
        raise ValueError('%s: bad NIfTI-1 header size' % path)
    magic = raw[344:348]
    if magic[:3] not in (b'n+1',):
        raise ValueError('%s: only single-file NIfTI-1 (magic n+1) is supported, got %r' % (path, magic))
    dim = struct.unpack(endian + '8h', raw[40:56])
    intent_code = struct.unpack(endian + 'h', raw[68:70])[0]
    datatype, bitpix = struct.unpack(endian + '2h', raw[70:74])
    pixdim = struct.unpack(endian + '8f', raw[76:108])
    vox_offset, slope, inter = struct.unpack(endian + '3f', raw[108:120])
    qform_code, sform_code = struct.unpack(endian + '2h', raw[252:256])
    qb, qc, qd, qx, qy, qz = struct.unpack(endian + '6f', raw[256:280])
    srow = np.array(struct.unpack(endian + '12f', raw[280:328]), dtype=np.float64).reshape(3, 4)
    if datatype not in _DTYPES:
        raise ValueError('%s: unsupported NIfTI datatype code %d' % (path, datatype))
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(endian)
    n = int(np.prod(shape)) if shape else 1
    off = int(vox_offset) if vox_offset >= 352 else 352
    data = np.frombuffer(raw, dtype=dt, count=n, offset=off).reshape(shape, order='F')
    data = data.astype(dt.newbyteorder('='), copy=True)
    if slope not in (0.0, 1.0) or inter != 0.0:
        if slope != 0.0 and not np.isnan(slope):
            data = data.astype(np.float64) * slope + inter
    if sform_code > 0:
        affine = np.vstack([srow, [0, 0, 0, 1]])
    elif qform_code > 0:
        affine = _quat_affine(qb, qc, qd, qx, qy, qz, pixdim[1], pixdim[2], pixdim[3], pixdim[0])
    else:
        affine = np.diag([pixdim[1] or 1.0, pixdim[2] or 1.0, pixdim[3] or 1.0, 1.0])
    if return_header:
        return data, affine, {'dim': dim, 'pixdim': pixdim, 'intent_code': intent_code, 'datatype': datatype,
                              'qform_code': qform_code, 'sform_code': sform_code}
    return data, affine


def save_nifti(array, path, affine=None, intent_code=0):
    """Write a little-endian single-file NIfTI-1 with an sform (and matching qform offsets)."""
    a = np.asarray(array)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype.name not in _CODES:
        a = a.astype(np.float32)
    if a.ndim > 7:
        raise ValueError('NIfTI-1 supports at most 7 dimensions')
    affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
    hdr = bytearray(348)
    struct.pack_into('<i', hdr, 0, 348)
    dim = [a.ndim] + list(a.shape) + [1] * (7 - a.ndim)
    struct.pack_into('<8h', hdr, 40, *dim)
    struct.pack_into('<h', hdr, 68, int(intent_code))
    struct.pack_into('<2h', hdr, 70, _CODES[a.dtype.name], a.dtype.itemsize * 8)
    zooms = np.sqrt((affine[:3, :3] ** 2).sum(0))
    pixdim = [1.0] + [float(z) for z in zooms] + [1.0] * 4
    struct.pack_into('<8f', hdr, 76, *pixdim)
    struct.pack_into('<3f', hdr, 108, 352.0, 1.0, 0.0)
    hdr[123] = 2                                           # xyzt_units: mm
    struct.pack_into('<2h', hdr, 252, 0, 2)                # qform_code 0, sform_code 2 (aligned)
    struct.pack_into('<12f', hdr, 280, *affine[:3, :].reshape(-1))
    hdr[344:348] = b'n+1\x00'
    with _open(path, 'wb') as f:
        f.write(bytes(hdr))
        f.write(b'\x00\x00\x00\x00')
        f.write(np.asfortranarray(a).astype(a.dtype.newbyteorder('<'), copy=False).tobytes(order='F'))


def aff2axcodes(affine):
    """Axis codes of an affine ('R','A','S' conventions), like nibabel.aff2axcodes
    (used for the RAI conversion of SCT warps: 3d_reg.py:399-417)."""
    R = np.asarray(affine, dtype=np.float64)[:3, :3]
    R = R / np.sqrt((R ** 2).sum(0, keepdims=True))
    labels = (('L', 'R'), ('P', 'A'), ('I', 'S'))
    codes = [None] * 3
    used = set()
    order = np.argsort(-np.abs(R).max(0))                  # assign the most axis-aligned columns first
    for col in order:
        rows = [r for r in np.argsort(-np.abs(R[:, col])) if r not in used]
        r = rows[0]
        used.add(r)
        codes[col] = labels[r][1] if R[r, col] > 0 else labels[r][0]
    return tuple(codes)
