"""TEST-ONLY stand-in: mmcv's `cache_randomness` only matters inside `cache_random_params` contexts (never used on the
CM-UNet data path), elsewhere it is the identity decorator."""


def cache_randomness(func):
    return func
