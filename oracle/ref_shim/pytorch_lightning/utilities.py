"""shim (test infrastructure)"""


def rank_zero_only(fn):
    return fn
