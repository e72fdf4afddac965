def ioff():
    pass


def __getattr__(name):
    def _missing(*a, **k):
        raise NotImplementedError(f"matplotlib stub: pyplot.{name}")
    return _missing
