"""TEST INFRASTRUCTURE — ctypes front-end of the CPU oracle (oracle/lbl_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (pyratbay_b200) never does.

Call signatures mirror the reference's CPython modules so tests read like the
reference's call sites:
  extinction(...)  <->  pyratbay/pyrat/extinction.py:197-208  (27 positional args)
  grid(...)        <->  pyratbay/pyrat/voigt.py:145-149
  interp_ec(...)   <->  pyratbay/opacity/line_sampling.py:451-456

`load_ref()` returns the *unmodified* reference modules compiled into oracle/_ref
(see oracle/build_ref.sh), or None if they have not been built.
"""
import ctypes
import importlib
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile liblbl_oracle.so (and oracle/_ref when /root/reference exists)."""
    so = os.path.join(_HERE, "liblbl_oracle.so")
    src = os.path.join(_HERE, "lbl_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liblbl_oracle.so"],
                              stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src_c"):
        ref_dir = os.path.join(_HERE, "_ref")
        have = (os.path.isdir(ref_dir)
                and any(f.startswith("_extcoeff") for f in os.listdir(ref_dir))
                and os.path.isdir(os.path.join(ref_dir, "shimmed", "pyratbay", "lib")))
        if force or not have:
            subprocess.check_call([os.path.join(_HERE, "build_ref.sh")],
                                  stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_voigt_grid.restype = ctypes.c_int
        _LIB.orc_extinction.restype = ctypes.c_int
        _LIB.orc_interp_ec.restype = ctypes.c_int
        _LIB.orc_interp_ec_mol.restype = ctypes.c_int
        _LIB.orc_plane_parallel_optical_depth.restype = ctypes.c_int
        _LIB.orc_transit_optical_depth.restype = ctypes.c_int
    return _LIB


def load_ref():
    """The reference's own _extcoeff / vprofile modules, or (None, None)."""
    ref_dir = os.path.join(_HERE, "_ref")
    if not os.path.isdir(ref_dir):
        return None, None
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        return importlib.import_module("_extcoeff"), importlib.import_module("vprofile")
    except ImportError:
        return None, None


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_ip)


def grid(profile, psize, index, lorentz, doppler, dwn, verb=0, quick_threshold=0):
    """In-place Voigt grid; same contract as vprofile.grid (vprofile.c:42-114)."""
    assert profile.dtype == np.float64 and profile.flags.c_contiguous
    assert psize.dtype == np.int64 and psize.flags.c_contiguous
    assert index.dtype == np.int64 and index.flags.c_contiguous
    lor, plor = _d(lorentz)
    dop, pdop = _d(doppler)
    ok = _lib().orc_voigt_grid(
        ctypes.c_int(len(lor)), ctypes.c_int(len(dop)), plor, pdop,
        ctypes.c_double(dwn), psize.ctypes.data_as(_ip), index.ctypes.data_as(_ip),
        profile.ctypes.data_as(_dp), ctypes.c_int(quick_threshold))
    if ok != 1:
        raise RuntimeError("oracle voigt grid failed")
    return 1


def extinction(ext, profile, psize, pindex, lorentz, doppler, wn, own, divisors,
               moldensity, molrad, molmass, isoimol, isomass, isoratio, isoz, isoiext,
               lwn, elow, gf, lID, cutoff, ethresh, temp, verb=0, add=0, resolution=0,
               counters=None):
    """In-place extinction for one (T,p); same contract as _extcoeff.extinction."""
    assert ext.dtype == np.float64 and ext.flags.c_contiguous and ext.ndim == 2
    keep = []

    def D(a):
        arr, p = _d(a)
        keep.append(arr)
        return p

    def I(a):
        arr, p = _i(a)
        keep.append(arr)
        return p

    cnt = np.zeros(4, np.int64)
    ok = _lib().orc_extinction(
        ext.ctypes.data_as(_dp), ctypes.c_int(ext.shape[0]), ctypes.c_int(ext.shape[1]),
        D(profile) if not (isinstance(profile, np.ndarray) and profile.dtype == np.float64
                           and profile.flags.c_contiguous) else profile.ctypes.data_as(_dp),
        I(psize), I(pindex),
        D(lorentz), ctypes.c_int(len(lorentz)), D(doppler), ctypes.c_int(len(doppler)),
        D(wn), D(own), ctypes.c_int64(len(own)),
        I(divisors), ctypes.c_int(len(divisors)),
        D(moldensity), D(molrad), D(molmass), ctypes.c_int(len(molmass)),
        I(isoimol), D(isomass), D(isoratio), D(isoz), I(isoiext),
        ctypes.c_int(len(isomass)),
        D(lwn), D(elow), D(gf), I(lID), ctypes.c_int64(len(lwn)),
        ctypes.c_double(cutoff), ctypes.c_double(ethresh), ctypes.c_double(temp),
        ctypes.c_int(int(add)), ctypes.c_int(int(resolution)),
        cnt.ctypes.data_as(_ip))
    if ok != 1:
        raise RuntimeError("oracle extinction failed")
    if counters is not None:
        counters[:] = cnt
    return 1


def _interp(fn, extinction_out, etable, ttable, temperatures, density, lay1, lay2):
    assert extinction_out.dtype == np.float64 and extinction_out.flags.c_contiguous
    etable = np.ascontiguousarray(etable, np.float64)
    nspec, ntemp, nlayers, nwave = etable.shape
    tt, ptt = _d(ttable)
    te, pte = _d(temperatures)
    de, pde = _d(density)
    fn(extinction_out.ctypes.data_as(_dp), etable.ctypes.data_as(_dp), ptt, pte, pde,
       ctypes.c_int(nspec), ctypes.c_int(ntemp), ctypes.c_int(nlayers),
       ctypes.c_int(nwave), ctypes.c_int(lay1), ctypes.c_int(lay2))
    return 1


def interp_ec(extinction_out, etable, ttable, temperatures, density, lay1, lay2):
    return _interp(_lib().orc_interp_ec, extinction_out, etable, ttable, temperatures,
                   density, lay1, lay2)


def interp_ec_per_mol(extinction_out, etable, ttable, temperatures, density, lay1, lay2):
    return _interp(_lib().orc_interp_ec_mol, extinction_out, etable, ttable, temperatures,
                   density, lay1, lay2)


def plane_parallel_optical_depth(depth, ideep, extinction, intervals, maxdepth, itop, ibottom):
    """In place, same contract as _trapezoid.plane_parallel_optical_depth
    (src_c/_trapezoid.c:147-211); ideep int32 [nwave]."""
    assert depth.dtype == np.float64 and depth.flags.c_contiguous
    assert ideep.dtype == np.int32 and ideep.flags.c_contiguous
    ext, pext = _d(extinction)
    dz, pdz = _d(intervals)
    nlayers, nwave = depth.shape
    _lib().orc_plane_parallel_optical_depth(
        depth.ctypes.data_as(_dp), ideep.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), pext,
        pdz, ctypes.c_double(maxdepth), ctypes.c_int(itop), ctypes.c_int(ibottom),
        ctypes.c_int(nlayers), ctypes.c_int(nwave))


def transit_optical_depth(depth, ideep, extinction, paths, maxdepth, itop, ibottom):
    """Slant optical depth of optic_depth.py:104-111 + _trapezoid.optdepth; `paths` is the
    [nlayers, nlayers] zero-padded matrix of ray paths (row r = raypath[r])."""
    assert depth.dtype == np.float64 and depth.flags.c_contiguous
    assert ideep.dtype == np.int32 and ideep.flags.c_contiguous
    ext, pext = _d(extinction)
    pa, ppa = _d(paths)
    nlayers, nwave = depth.shape
    _lib().orc_transit_optical_depth(
        depth.ctypes.data_as(_dp), ideep.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), pext,
        ppa, ctypes.c_double(maxdepth), ctypes.c_int(itop), ctypes.c_int(ibottom),
        ctypes.c_int(nlayers), ctypes.c_int(nwave))
