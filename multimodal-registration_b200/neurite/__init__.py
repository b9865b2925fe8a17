"""Mirror of the neurite names on the reference's deformation hot path."""
from . import utils   # noqa: F401
