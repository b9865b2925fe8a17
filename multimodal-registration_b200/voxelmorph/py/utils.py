"""vxm.py.utils mirror (volume file I/O used by the reference scripts).  NIfTI support is a
"next" row; .npy/.npz volumes work today."""
import numpy as np


def load_volfile(filename, np_var='vol', add_batch_axis=False, add_feat_axis=False,
                 pad_shape=None, resize_factor=1, ret_affine=False):
    if filename.endswith('.npy'):
        vol, affine = np.load(filename), np.eye(4)
    elif filename.endswith('.npz'):
        npz = np.load(filename)
        vol = npz[np_var] if np_var in npz else next(iter(npz.values()))
        affine = npz['affine'] if 'affine' in npz else np.eye(4)
    elif filename.endswith(('.nii', '.nii.gz')):
        from ..._nifti import load_nifti
        vol, affine = load_nifti(filename)
        vol = vol.squeeze()
    else:
        raise ValueError('unknown filetype for %s' % filename)
    if add_feat_axis:
        vol = vol[..., np.newaxis]
    if add_batch_axis:
        vol = vol[np.newaxis, ...]
    return (vol, affine) if ret_affine else vol


def save_volfile(array, filename, affine=None):
    if filename.endswith(('.nii', '.nii.gz')):
        from ..._nifti import save_nifti
        save_nifti(array, filename, affine)
    elif filename.endswith('.npz'):
        np.savez_compressed(filename, vol=array, affine=np.eye(4) if affine is None else affine)
    elif filename.endswith('.npy'):
        np.save(filename, array)
    else:
        raise ValueError('unknown filetype for %s' % filename)
