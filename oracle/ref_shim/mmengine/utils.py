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


def is_seq_of(seq, expected_type, seq_type=None):
    import collections.abc as abc
    if not isinstance(seq, seq_type or abc.Sequence):
        return False
    return all(isinstance(x, expected_type) for x in seq)


def is_list_of(seq, expected_type):
    return is_seq_of(seq, expected_type, seq_type=list)
