"""vxm.tf.utils mirror: setup_device (train_synthmorph.py:192)."""
import os


def setup_device(gpuid=None):
    """Returns (device string, number of devices) like voxelmorph: ``len(gpuid.split(','))``.
    One process drives one GPU here, so with torchrun the local rank picks the device."""
    import torch
    if gpuid is not None and not isinstance(gpuid, str):
        gpuid = str(gpuid)
    if gpuid is None or gpuid == '-1':
        raise RuntimeError('setup_device: CPU execution requested, but the deformation engine has no CPU path')
    ids = gpuid.split(',')
    nb_devices = len(ids)
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dev = int(ids[local % nb_devices])
    torch.cuda.set_device(dev)
    return 'cuda:%d' % dev, nb_devices
