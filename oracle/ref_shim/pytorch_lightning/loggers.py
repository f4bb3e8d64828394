"""shim (test infrastructure)"""


class WandbLogger:
    def __init__(self, *a, **k):
        pass
