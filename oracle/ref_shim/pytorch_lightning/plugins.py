"""shim (test infrastructure): plugin classes only used in isinstance() checks"""


class DDPPlugin:
    pass


class DDP2Plugin:
    pass
