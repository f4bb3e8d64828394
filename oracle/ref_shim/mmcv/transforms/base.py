"""TEST-ONLY stand-in for mmcv.transforms.base.BaseTransform: `__call__` delegates to `transform`."""


class BaseTransform:
    def __call__(self, results):
        return self.transform(results)

    def transform(self, results):
        raise NotImplementedError
