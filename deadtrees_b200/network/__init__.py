from .segmodel import SemSegment  # noqa: F401
from .unet import Unet  # noqa: F401
