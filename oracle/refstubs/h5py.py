"""Stand-in for h5py (absent); only petitRADTRANS tables would need it."""


class File:
    def __init__(self, *a, **k):
        raise NotImplementedError("h5py stub")
