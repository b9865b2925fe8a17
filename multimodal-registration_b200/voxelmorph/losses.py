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
