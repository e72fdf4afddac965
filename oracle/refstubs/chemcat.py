"""Stand-in for chemcat (absent); equilibrium chemistry is never used here."""


class Network:
    def __init__(self, *a, **k):
        raise NotImplementedError("chemcat stub")


def __getattr__(name):
    raise AttributeError(name)
