"""`python -m contrastive_masked_unet_b200.build [--force] [-v]` -- builds libcmu_b200.so (see `_buildlib`).

Python binds an imported submodule as an attribute of its package, which would replace the package-level model factory
`build(cfg)` (modules.build, the stand-in for `MODELS.build`) by this module as soon as somebody imports
`contrastive_masked_unet_b200.build`.  To keep `pkg.build(cfg)` working either way, this module is callable and
forwards to the factory."""
import sys
import types

from ._buildlib import LIB, build as build_library, sources  # noqa: F401


class _CallableModule(types.ModuleType):
    def __call__(self, cfg):
        from .modules import build as build_model
        return build_model(cfg)


sys.modules[__name__].__class__ = _CallableModule
build = build_library          # `from contrastive_masked_unet_b200.build import build` keeps meaning "build the library"

if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
