"""shim (test infrastructure)"""


class ModelCheckpoint:
    def __init__(self, *a, **k):
        pass
