def __getattr__(name):
    def _missing(*a, **k):
        raise NotImplementedError(f"mc3 stub: stats.{name}")
    return _missing
