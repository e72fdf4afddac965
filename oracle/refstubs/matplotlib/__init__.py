"""Import-time stand-in for matplotlib (absent here); plotting is never called."""
from . import colors, pyplot  # noqa: F401

rcParams = {}


def use(*a, **k):
    pass


def __getattr__(name):
    import types
    mod = types.ModuleType(name)
    return mod
