"""TEST-ONLY stand-in for the un-vendored `mmengine==0.10.5` dependency of the reference
(Pretraining/CM-UNet/environment.yml:21). It exists so that the UNMODIFIED reference sources
under /root/reference can be imported in the build container to mint golden vectors
(oracle/make_goldens.py). It is never imported by the product package."""
__version__ = '0.10.5'
from . import utils, registry, model, dist  # noqa: F401
