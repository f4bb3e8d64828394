"""TEST-ONLY stand-in for `mmcv==2.2.0` (Pretraining/CM-UNet/environment.yml:22)."""
__version__ = '2.2.0'
from . import cnn  # noqa: F401
