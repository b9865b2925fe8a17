"""Mirror of the voxelmorph losses the reference's training script uses (train_synthmorph.py:301-306):
``Dice().loss(y_true, y_pred)`` and ``Grad('l2', loss_mult).loss(None, flow)``.  The voxel-level sums
and gradients run in libdfm (dfm_dice_*, dfm_grad_l2_*); other losses / penalties are outside the path."""
import torch

from .. import _host, ops


class Dice:
    """N-D Dice for one-hot / probabilistic maps [B, *vol, C]; ``loss`` returns -mean Dice."""

    def loss(self, y_true, y_pred):
        return ops.dice_loss(_host.to_device(y_true, torch.float32, tag='y_true'),
                             _host.to_device(y_pred, torch.float32, tag='y_pred'))


class Grad:
    """Spatial gradient penalty of a displacement field.  Only ``penalty='l2'`` is on the path."""

    def __init__(self, penalty='l1', loss_mult=None, vox_weight=None):
        if penalty != 'l2':
            raise NotImplementedError("Grad: only penalty='l2' (train_synthmorph.py:306) is implemented")
        if vox_weight is not None:
            raise NotImplementedError('Grad: vox_weight is outside the hot path')
        self.penalty = penalty
        self.loss_mult = loss_mult

    def loss(self, _, y_pred):
        mult = 1.0 if self.loss_mult is None else float(self.loss_mult)
        return ops.grad_l2_loss(_host.to_device(y_pred, torch.float32, tag='flow'), mult)


def dice_loss_zeropad(y_true, y_pred):
    """The reference's own ``losses.dice_loss_zeropad`` (losses.py:11-69) *as documented*: Dice over the labels
    1..C-1 of batch item 0, ignoring every voxel where the label-0 channel of either map is >= 1 (the
    zero-padded regions); -mean(divide_no_nan(top, bottom)).  (The shipped function raises unconditionally --
    its ``raise ValueError(err)`` at losses.py:32 is not indented under the ``if ndims != 3`` -- so this follows
    the docstring and the code after that line.)  The masking is element-wise torch on the device, the sums and
    the gradient are libdfm's Dice kernels."""
    yt = _host.to_device(y_true, torch.float32, tag='y_true')
    yp = _host.to_device(y_pred, torch.float32, tag='y_pred')
    if yp.dim() != 5:
        raise ValueError('The Dice loss computed only on regions with no zero-padding can only be used on 3D volumes '
                         'but the dimension of the object is: %d. The expected input should be of shape '
                         '[None, x, y, z, n_labels] but received: %s and %s' % (yp.dim() - 2, list(yt.shape), list(yp.shape)))
    keep = ~((yt[0, ..., 0] >= 1) | (yp[0, ..., 0] >= 1))                     # losses.py:38-43
    m = keep.to(yp.dtype)[None, ..., None]
    yt0 = (yt[0:1, ..., 1:] * m).contiguous()                                  # labels 1.. of item 0, masked (:50-66)
    yp0 = (yp[0:1, ..., 1:] * m).contiguous()
    if yp0.shape[-1] == 0:
        return yp.sum() * 0.0
    return ops.dice_loss(yt0, yp0)
