"""`pyratbay.lib._extcoeff` on the GPU engine (signatures of src_c/_extcoeff.c)."""
import numpy as np

from . import client

__all__ = ["extinction", "interp_ec", "interp_ec_per_mol"]


def extinction(ext, profile, psize, pindex, lorentz, doppler, wn, own, divisors, moldensity,
               molrad, molmass, isoimol, isomass, isoratio, isoz, isoiext, lwn, elow, gf, lid,
               cutoff, ethresh, temp, verb=0, add=0, resolution=0):
    """ec.extinction(...) (_extcoeff.c:87-345): adds the extinction of one (T,p) unit to `ext`
    [nextinct, nwave] in place; returns 1."""
    if not (isinstance(ext, np.ndarray) and ext.dtype == np.float64 and ext.ndim == 2
            and ext.flags.c_contiguous):
        raise TypeError("extinction: 'ext' must be a C-contiguous float64 [nextinct, nwave] array")
    static = {name: client.static_key(np.asarray(arr, dt)) for name, arr, dt in (
        ("profile", profile, np.float64), ("psize", psize, np.int64),
        ("pindex", pindex, np.int64), ("lorentz", lorentz, np.float64),
        ("doppler", doppler, np.float64), ("wn", wn, np.float64), ("own", own, np.float64),
        ("divisors", divisors, np.int64), ("molrad", molrad, np.float64),
        ("molmass", molmass, np.float64), ("isoimol", isoimol, np.int64),
        ("isomass", isomass, np.float64), ("isoratio", isoratio, np.float64),
        ("lwn", lwn, np.float64), ("elow", elow, np.float64), ("gf", gf, np.float64),
        ("lid", lid, np.int64))}
    unit = dict(moldensity=np.asarray(moldensity, np.float64),
                isoz=np.asarray(isoz, np.float64), isoiext=np.asarray(isoiext, np.int64),
                cutoff=float(cutoff), ethresh=float(ethresh), temp=float(temp), add=int(add),
                resolution=int(resolution), nextinct=int(ext.shape[0]))
    got = client.request("extinction", static, unit)
    rows = got.shape[0]
    ext[:rows] += got
    return 1


def _interp(op, extinction_out, etable, ttable, temperatures, density, lay1, lay2):
    if not (isinstance(extinction_out, np.ndarray) and extinction_out.dtype == np.float64
            and extinction_out.flags.c_contiguous):
        raise TypeError("interp_ec: 'extinction' must be a C-contiguous float64 array")
    key = client.static_key(np.asarray(etable, np.float64))
    got = client.request(op, key, np.asarray(ttable, np.float64),
                         np.asarray(temperatures, np.float64), np.asarray(density, np.float64),
                         int(lay1), int(lay2), extinction_out)
    extinction_out[...] = got
    return 1


def interp_ec(extinction, etable, ttable, temperatures, density, lay1, lay2):
    """ec.interp_ec (_extcoeff.c:367-418): accumulates into extinction [nlayers, nwave]."""
    return _interp("interp_ec", extinction, etable, ttable, temperatures, density, lay1, lay2)


def interp_ec_per_mol(extinction, etable, ttable, temperatures, density, lay1, lay2):
    """ec.interp_ec_per_mol (_extcoeff.c:421-472): extinction [nspec, nlayers, nwave]."""
    return _interp("interp_ec_per_mol", extinction, etable, ttable, temperatures, density, lay1,
                   lay2)
