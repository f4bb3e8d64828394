"""TEST-ONLY stand-ins for the parts of `mmcv.transforms` (mmcv 2.2.0, un-vendored) that the reference's data pipeline
files import at module level (cmae/datasets/pipelines/{processing,auto_augment,formatting,wrappers}.py)."""
from .base import BaseTransform  # noqa: F401


class Compose:
    def __init__(self, transforms=None):
        self.transforms = list(transforms or [])

    def __call__(self, data):
        for t in self.transforms:
            data = t(data)
            if data is None:
                return None
        return data


class RandomChoice(BaseTransform):
    def __init__(self, transforms=None, prob=None):
        self.transforms, self.prob = transforms, prob

    def transform(self, results):
        raise NotImplementedError('RandomChoice is not on the CM-UNet data path')
