"""multimodal-registration_b200: B200-native deformation engine behind the voxelmorph /
neurite API used by ivadomed/multimodal-registration.

  ops         device-level functional ops on torch CUDA tensors (libdfm.so via ctypes)
  voxelmorph  mirror of the voxelmorph names the reference scripts call
              (layers.SpatialTransformer/VecInt/RescaleTransform, utils.transform/compose/
              rescale_dense_transform/integrate_vec, networks.Transform/VxmDense, py.utils)
  neurite     mirror of neurite.utils.interpn/resize/zoom and utils.augment.draw_perlin
  sct_warp    the SCT warp-file convention of 3d_reg.py:390-422 (rescale, RAI components, intent 1007)
  metrics     NMI joint histogram, zero-padding box and segmentation-overlap metrics of the eval scripts
  pipelines   device-resident tails of the registration scripts (two-step cascade, sub-volume stitching, export)

The directory name is not a Python identifier; import it through the top-level alias module
``multimodal_registration_b200`` (repo root), or call ``install_shims()`` to register the
mirrors as ``voxelmorph`` and ``neurite`` so ``import voxelmorph as vxm`` resolves to them.
"""
import sys

from . import _lib, ops, sharding          # noqa: F401
from . import neurite, voxelmorph   # noqa: F401
from . import sct_warp             # noqa: F401
from . import pipelines            # noqa: F401
from . import metrics              # noqa: F401

__version__ = '0.1.0'


def install_shims():
    """Register the mirrors under the names the reference scripts import."""
    sys.modules.setdefault('voxelmorph', voxelmorph)
    sys.modules.setdefault('neurite', neurite)
    for mod in (voxelmorph, neurite):
        prefix = mod.__name__
        for name, sub in list(sys.modules.items()):
            if name.startswith(prefix + '.') and sub is not None:
                sys.modules.setdefault(mod.__name__.rsplit('.', 1)[-1] + name[len(prefix):], sub)
    return voxelmorph, neurite
