"""Mirror of the voxelmorph names on the reference's deformation hot path."""
from . import layers, losses, networks, utils   # noqa: F401
from . import py, tf                    # noqa: F401
