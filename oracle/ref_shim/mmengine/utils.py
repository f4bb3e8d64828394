"""Subset of mmengine.utils used by cmae/__init__.py:5 (test-only shim)."""


def is_str(x):
    return isinstance(x, str)


def digit_version(version_str, length=4):
    parts = []
    for p in version_str.split('+')[0].split('.'):
        digits = ''.join(ch for ch in p if ch.isdigit())
        parts.append(int(digits) if digits else 0)
    parts = (parts + [0] * length)[:length]
    return tuple(parts)
