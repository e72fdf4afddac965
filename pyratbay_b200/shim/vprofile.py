"""`pyratbay.lib.vprofile` on the GPU engine (signature of src_c/vprofile.c:53-57)."""
import numpy as np

from . import client

__all__ = ["grid"]


def grid(profile, psize, index, lorentz, doppler, dwn, verb=0):
    """vp.grid(...) (vprofile.c:42-114): fills `profile`, `psize`, `index` in place; returns 1."""
    for name, arr, dt in (("profile", profile, np.float64), ("psize", psize, np.int64),
                          ("index", index, np.int64)):
        if not (isinstance(arr, np.ndarray) and arr.dtype == dt and arr.flags.c_contiguous):
            raise TypeError(f"grid: '{name}' must be a C-contiguous {np.dtype(dt).name} array")
    size, idx = client.request("grid_sizes", np.asarray(lorentz, np.float64),
                               np.asarray(doppler, np.float64), float(dwn), psize, profile.size)
    psize[...] = size
    index[...] = idx
    profile[...] = client.request("grid_profile")
    return 1
