"""TEST-ONLY stand-in for `mmcv==2.2.0` (Pretraining/CM-UNet/environment.yml:22)."""
__version__ = '2.2.0'
from . import cnn  # noqa: F401
from . import transforms  # noqa: F401,E402


def imcrop(img, bboxes, scale=1.0, pad_fill=None):
    """[mmcv, from memory] single bbox (x1, y1, x2, y2), inclusive corners, clipped to the image."""
    import numpy as np
    x1, y1, x2, y2 = [int(v) for v in np.asarray(bboxes).reshape(-1)[:4]]
    h, w = img.shape[:2]
    x1, y1, x2, y2 = max(x1, 0), max(y1, 0), min(x2, w - 1), min(y2, h - 1)
    return img[y1:y2 + 1, x1:x2 + 1, ...]


def imflip(img, direction='horizontal'):
    import numpy as np
    return np.flip(img, axis=1) if direction == 'horizontal' else np.flip(img, axis=0)
