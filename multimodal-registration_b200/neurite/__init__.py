"""Mirror of the neurite names on the reference's deformation hot path."""
from . import utils   # noqa: F401
from . import models  # noqa: F401
