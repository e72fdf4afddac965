"""Python handle over the C-ABI engine (include/pb200_lbl.h).

`Engine` is deliberately thin: it marshals NumPy arrays to the plain-pointer ABI and
nothing else.  The reference-shaped objects (Voigt, Line_By_Line, extinction.*) sit on top.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import c_double_p, c_int64_p, check


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int64_p)


def voigt_grid(profile, psize, index, lorentz, doppler, dwn, verb=0, device=0):
    """Drop-in for lib.vprofile.grid (src_c/vprofile.c:42-114): fills `profile`, `psize`
    and `index` in place, returns 1.  Computed on the GPU."""
    lib = _lib.load()
    _lib.require_device()
    for name, arr, dt in (("profile", profile, np.float64), ("psize", psize, np.int64),
                          ("index", index, np.int64)):
        if not (isinstance(arr, np.ndarray) and arr.dtype == dt and arr.flags.c_contiguous):
            raise TypeError(f"voigt_grid: '{name}' must be a C-contiguous {dt.__name__} array")
    lor, dop = _f64(lorentz), _f64(doppler)
    if psize.shape != (len(lor), len(dop)) or index.shape != psize.shape:
        raise ValueError("voigt_grid: psize/index must have shape [nlor, ndop]")
    check(lib.pb200_voigt_grid(
        ctypes.c_int(device), ctypes.c_int(len(lor)), ctypes.c_int(len(dop)), _dp(lor),
        _dp(dop), ctypes.c_double(dwn), _ip(psize), _ip(index), _dp(profile),
        ctypes.c_int64(profile.size)))
    return 1


def interp_ec(extinction, etable, ttable, temperatures, density, lay1, lay2, device=0):
    """Drop-in for lib._extcoeff.interp_ec (src_c/_extcoeff.c:367-418); `extinction`
    [nlayers, nwave] is accumulated in place."""
    return _interp(False, extinction, etable, ttable, temperatures, density, lay1, lay2,
                   device)


def interp_ec_per_mol(extinction, etable, ttable, temperatures, density, lay1, lay2,
                      device=0):
    """Drop-in for lib._extcoeff.interp_ec_per_mol (src_c/_extcoeff.c:421-472)."""
    return _interp(True, extinction, etable, ttable, temperatures, density, lay1, lay2,
                   device)


def _interp(per_mol, extinction, etable, ttable, temperatures, density, lay1, lay2, device):
    lib = _lib.load()
    _lib.require_device()
    if not (isinstance(extinction, np.ndarray) and extinction.dtype == np.float64
            and extinction.flags.c_contiguous):
        raise TypeError("interp_ec: 'extinction' must be a C-contiguous float64 array")
    etable = _f64(etable)
    if etable.ndim != 4:
        raise ValueError("interp_ec: etable must be [nspec, ntemp, nlayers, nwave]")
    nspec, ntemp, nlayers, nwave = etable.shape
    want = (nspec, nlayers, nwave) if per_mol else (nlayers, nwave)
    if extinction.shape != want:
        raise ValueError(f"interp_ec: extinction must have shape {want}")
    tt, te, de = _f64(ttable), _f64(temperatures), _f64(density)
    if tt.shape != (ntemp,) or te.shape != (nlayers,) or de.shape != (nlayers, nspec):
        raise ValueError("interp_ec: inconsistent ttable/temperatures/density shapes")
    fn = lib.pb200_interp_ec_per_mol if per_mol else lib.pb200_interp_ec
    check(fn(ctypes.c_int(device), _dp(extinction), _dp(etable), _dp(tt), _dp(te), _dp(de),
             ctypes.c_int(nspec), ctypes.c_int(ntemp), ctypes.c_int(nlayers),
             ctypes.c_int(nwave), ctypes.c_int(int(lay1)), ctypes.c_int(int(lay2))))
    return 1


def interp_ec_device(ext_ptr, etable_ptr, ttable, temperatures, density, shape, lay1, lay2,
                     per_mol=False, device=0, stream=0):
    """Device-resident form: `ext_ptr` / `etable_ptr` are device addresses (e.g. torch
    tensors' data_ptr()), `shape` = (nspec, ntemp, nlayers, nwave)."""
    lib = _lib.load()
    nspec, ntemp, nlayers, nwave = (int(v) for v in shape)
    tt, te, de = _f64(ttable), _f64(temperatures), _f64(density)
    check(lib.pb200_interp_ec_dev(
        ctypes.c_int(device), ctypes.c_void_p(ext_ptr), ctypes.c_void_p(etable_ptr), _dp(tt),
        _dp(te), _dp(de), ctypes.c_int(nspec), ctypes.c_int(ntemp), ctypes.c_int(nlayers),
        ctypes.c_int(nwave), ctypes.c_int(int(lay1)), ctypes.c_int(int(lay2)),
        ctypes.c_int(int(per_mol)), ctypes.c_void_p(stream)))
    return 1


class Engine:
    """Owns one pb200_engine handle (device copies of grids, Voigt table and lines)."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        _lib.require_device()
        self.device = int(device)
        handle = ctypes.c_void_p()
        check(self._lib.pb200_engine_create(ctypes.c_int(self.device), ctypes.byref(handle)))
        self._h = handle
        self.nwave = 0
        self.nmol = 0
        self.niso = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pb200_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- static inputs ---------------------------------------------------------------
    def set_grid(self, wn, own, divisors):
        wn, own, div = _f64(wn), _f64(own), _i64(divisors)
        check(self._lib.pb200_engine_set_grid(
            self._h, _dp(wn), ctypes.c_int64(len(wn)), _dp(own), ctypes.c_int64(len(own)),
            _ip(div), ctypes.c_int(len(div))))
        self.nwave = len(wn)

    def build_voigt(self, lorentz, doppler, dwn, psize, pindex, cutoff):
        """psize [nlor,ndop] int64 is updated in place, pindex is filled."""
        lor, dop = _f64(lorentz), _f64(doppler)
        assert psize.dtype == np.int64 and psize.flags.c_contiguous
        assert pindex.dtype == np.int64 and pindex.flags.c_contiguous
        check(self._lib.pb200_engine_build_voigt(
            self._h, ctypes.c_int(len(lor)), ctypes.c_int(len(dop)), _dp(lor), _dp(dop),
            ctypes.c_double(dwn), _ip(psize), _ip(pindex), ctypes.c_double(cutoff)))

    def set_voigt(self, lorentz, doppler, psize, pindex, profile, cutoff):
        lor, dop = _f64(lorentz), _f64(doppler)
        ps, pi, pr = _i64(psize), _i64(pindex), _f64(profile)
        check(self._lib.pb200_engine_set_voigt(
            self._h, ctypes.c_int(len(lor)), ctypes.c_int(len(dop)), _dp(lor), _dp(dop),
            _ip(ps), _ip(pi), _dp(pr), ctypes.c_int64(pr.size), ctypes.c_double(cutoff)))

    def profile_len(self):
        return int(self._lib.pb200_engine_profile_len(self._h))

    def get_profile(self, out=None):
        n = self.profile_len()
        if out is None:
            out = np.empty(n, np.float64)
        check(self._lib.pb200_engine_get_profile(self._h, _dp(out), ctypes.c_int64(out.size)))
        return out

    def set_species(self, mol_radius, mol_mass, iso_imol, iso_mass, iso_ratio):
        mr, mm = _f64(mol_radius), _f64(mol_mass)
        ii, im, ir = _i64(iso_imol), _f64(iso_mass), _f64(iso_ratio)
        check(self._lib.pb200_engine_set_species(
            self._h, ctypes.c_int(len(mm)), _dp(mr), _dp(mm), ctypes.c_int(len(im)), _ip(ii),
            _dp(im), _dp(ir)))
        self.nmol, self.niso = len(mm), len(im)

    def set_partition(self, temp, z):
        t, zz = _f64(temp), _f64(z)
        if zz.shape != (self.niso, len(t)):
            raise ValueError("set_partition: z must be [niso, ntemp]")
        check(self._lib.pb200_engine_set_partition(self._h, ctypes.c_int(len(t)), _dp(t),
                                                   _dp(zz)))

    def set_lines(self, wn, elow, gf, isoid):
        w, e, g, i = _f64(wn), _f64(elow), _f64(gf), _i64(isoid)
        if not (len(w) == len(e) == len(g) == len(i)):
            raise ValueError("set_lines: arrays differ in length")
        check(self._lib.pb200_engine_set_lines(self._h, ctypes.c_int64(len(w)), _dp(w), _dp(e),
                                               _dp(g), _ip(i)))

    def line_stats(self):
        s = np.zeros(3, np.int64)
        check(self._lib.pb200_engine_line_stats(self._h, _ip(s)))
        return {"in_window": int(s[0]), "groups": int(s[1]), "nadd": int(s[2])}

    # -- batched extinction ----------------------------------------------------------
    def extinction_batch(self, temp, density, isoz, iso_iext, nextinct, ethresh, add,
                         resolution, out=None, counters=False, out_device_ptr=None,
                         stream=0):
        """Evaluate len(temp) (T,p) units.  Returns `out` [n_units, nrows, nwave] (host) or
        None when `out_device_ptr` is given; with counters=True returns (out, counters)."""
        t = _f64(temp)
        n_units = len(t)
        d = _f64(density).reshape(n_units, self.nmol)
        z = None if isoz is None else _f64(isoz).reshape(n_units, self.niso)
        ie = _i64(iso_iext)
        nrows = 1 if add else int(nextinct)
        cnt = np.zeros((n_units, 6), np.int64) if counters else None
        cnt_p = _ip(cnt) if counters else None
        z_p = _dp(z) if z is not None else None
        if out_device_ptr is not None:
            check(self._lib.pb200_extinction_batch_dev(
                self._h, ctypes.c_int(n_units), _dp(t), _dp(d), z_p, _ip(ie),
                ctypes.c_int(int(nextinct)), ctypes.c_double(ethresh), ctypes.c_int(int(add)),
                ctypes.c_int(int(resolution)), ctypes.c_void_p(out_device_ptr), cnt_p,
                ctypes.c_void_p(stream)))
            return (None, cnt) if counters else None
        if out is None:
            out = np.empty((n_units, nrows, self.nwave), np.float64)
        elif not (out.dtype == np.float64 and out.flags.c_contiguous
                  and out.size == n_units * nrows * self.nwave):
            raise ValueError("extinction_batch: bad output array")
        check(self._lib.pb200_extinction_batch_host(
            self._h, ctypes.c_int(n_units), _dp(t), _dp(d), z_p, _ip(ie),
            ctypes.c_int(int(nextinct)), ctypes.c_double(ethresh), ctypes.c_int(int(add)),
            ctypes.c_int(int(resolution)), _dp(out), cnt_p))
        return (out, cnt) if counters else out

    def extinction_batch_ptr(self, n_units, temp_ptr, density_ptr, isoz_ptr, iso_iext,
                             nextinct, ethresh, add, resolution, out_ptr, out_on_device):
        """Raw-address form for pinned host buffers (bench e2e path): all *_ptr are integer
        addresses of float64 buffers with the documented shapes."""
        ie = _i64(iso_iext)
        fn = (self._lib.pb200_extinction_batch_dev if out_on_device
              else self._lib.pb200_extinction_batch_host)
        args = [self._h, ctypes.c_int(n_units), ctypes.c_void_p(temp_ptr),
                ctypes.c_void_p(density_ptr), ctypes.c_void_p(isoz_ptr), _ip(ie),
                ctypes.c_int(int(nextinct)), ctypes.c_double(ethresh), ctypes.c_int(int(add)),
                ctypes.c_int(int(resolution)), ctypes.c_void_p(out_ptr), None]
        if out_on_device:
            args.append(ctypes.c_void_p(0))
        check(fn(*args))

    def last_timing(self):
        ms = np.zeros(5, np.float64)
        check(self._lib.pb200_engine_last_timing(self._h, _dp(ms)))
        return {"strengths_ms": ms[0], "accumulate_ms": ms[1], "h2d_ms": ms[2],
                "d2h_ms": ms[3], "total_ms": ms[4], "dense_ms": self.dense_ms(),
                "dense_units": self.dense_units()}

    def launch_count(self):
        return int(self._lib.pb200_engine_launch_count(self._h))

    def dense_units(self):
        """(unit, isotope) pairs of the last batch evaluated by the dense-convolution kernel."""
        return int(self._lib.pb200_engine_dense_units(self._h))

    def dense_ms(self):
        """Device time (ms) of the dense-convolution kernels in the last batch."""
        return float(self._lib.pb200_engine_dense_ms(self._h))

    def stream_ptr(self):
        """Address of the engine's cudaStream_t (for torch.cuda.ExternalStream)."""
        return int(self._lib.pb200_engine_stream(self._h) or 0)


class DeviceTable:
    """Handle over a cross-section table [nspec, ntemp, nlayers, nwave] resident in HBM
    (pb200_table_*): temperature interpolation without per-call allocation or synchronisation.
    `table` is anything with data_ptr() (a torch CUDA tensor); it is kept alive here."""

    def __init__(self, table, ttable, device=0):
        self._lib = _lib.load()
        _lib.require_device()
        self.table = table
        self.shape = tuple(int(v) for v in table.shape)
        if len(self.shape) != 4:
            raise ValueError("DeviceTable: table must be [nspec, ntemp, nlayers, nwave]")
        nspec, ntemp, nlayers, nwave = self.shape
        tt = _f64(ttable)
        if tt.shape != (ntemp,):
            raise ValueError("DeviceTable: ttable must have ntemp samples")
        handle = ctypes.c_void_p()
        check(self._lib.pb200_table_create(
            ctypes.c_int(int(device)), ctypes.c_void_p(table.data_ptr()), _dp(tt),
            ctypes.c_int(nspec), ctypes.c_int(ntemp), ctypes.c_int(nlayers), ctypes.c_int(nwave),
            ctypes.byref(handle)))
        self._h = handle
        self._fn = self._lib.pb200_table_interp

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pb200_table_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def interp(self, temperature, density, lay1, lay2, per_mol, ext_ptr, overwrite=True,
               stream=0, sync=False):
        """Queue interp_ec / interp_ec_per_mol into the device buffer at `ext_ptr`
        ([nlayers, nwave] or [nspec, nlayers, nwave]).  Returns without waiting unless sync."""
        nspec, _, nlayers, _ = self.shape
        te = np.ascontiguousarray(temperature, np.float64)
        de = np.ascontiguousarray(density, np.float64)
        if te.size != nlayers or de.size != nlayers * nspec:
            raise ValueError("DeviceTable.interp: inconsistent temperature/density shapes")
        status = self._fn(self._h, te.ctypes.data_as(c_double_p), de.ctypes.data_as(c_double_p),
                          ctypes.c_int(int(lay1)), ctypes.c_int(int(lay2)),
                          ctypes.c_int(int(per_mol)), ctypes.c_void_p(ext_ptr),
                          ctypes.c_int(int(overwrite)), ctypes.c_void_p(stream),
                          ctypes.c_int(int(sync)))
        if status:
            check(status)

    def launch_count(self):
        return int(self._lib.pb200_table_launch_count(self._h))


def regrid_table_device(table_ptr, shape, t_brackets, p_brackets, wave_idx, out_ptr, take_log,
                        accumulate=False, device=0, stream=0):
    """pb200_regrid_table_dev: `shape` = (ntemp, nlayers, nwave) of the source table on the
    device; *_brackets = (lo, hi, weight) arrays per output sample; wave_idx int32 or None."""
    lib = _lib.load()
    ntemp, nlayers, nwave = (int(v) for v in shape)
    c_int_p = ctypes.POINTER(ctypes.c_int)

    def ints(a):
        a = np.ascontiguousarray(a, np.int32)
        return a, a.ctypes.data_as(c_int_p)
    tlo, ptlo = ints(t_brackets[0])
    thi, pthi = ints(t_brackets[1])
    tf = _f64(t_brackets[2])
    plo, pplo = ints(p_brackets[0])
    phi, pphi = ints(p_brackets[1])
    pf = _f64(p_brackets[2])
    if wave_idx is None:
        widx, pw, nwave_out = None, None, nwave
    else:
        widx, pw = ints(wave_idx)
        nwave_out = len(widx)
    check(lib.pb200_regrid_table_dev(
        ctypes.c_int(int(device)), ctypes.c_void_p(table_ptr), ctypes.c_int(ntemp),
        ctypes.c_int(nlayers), ctypes.c_int(nwave), ptlo, pthi, _dp(tf), ctypes.c_int(len(tlo)),
        pplo, pphi, _dp(pf), ctypes.c_int(len(plo)), pw, ctypes.c_int(nwave_out),
        ctypes.c_int(int(take_log)), ctypes.c_void_p(out_ptr), ctypes.c_int(int(accumulate)),
        ctypes.c_void_p(stream)))
    return (len(tlo), len(plo), nwave_out)


def device_ceilings(device=0, l2_mbytes=32, reps=3):
    """Measured fp64-FMA TFLOP/s and L2-resident read GB/s (bench.py roofline denominators)."""
    lib = _lib.load()
    _lib.require_device()
    tf, gb = ctypes.c_double(0.0), ctypes.c_double(0.0)
    check(lib.pb200_bench_fp64(ctypes.c_int(device), ctypes.c_int(reps), ctypes.byref(tf)))
    check(lib.pb200_bench_l2(ctypes.c_int(device), ctypes.c_int(l2_mbytes), ctypes.c_int(reps),
                             ctypes.byref(gb)))
    return tf.value, gb.value


def nearest_thresholds(grid):
    """thr[j] = smallest double whose nearest sample of the increasing `grid` is >= j
    (pb200_nearest_thresholds; host only)."""
    g = _f64(grid)
    thr = np.zeros(len(g), np.float64)
    check(_lib.load().pb200_nearest_thresholds(_dp(g), ctypes.c_int(len(g)), _dp(thr)))
    return thr


def selftest_exact(steps, grid, n=1 << 24, seed=1, device=0):
    """(quotient mismatches, nearest-index mismatches) of the device self-test
    (pb200_selftest_exact): both must be zero."""
    lib = _lib.load()
    _lib.require_device()
    s, g = _f64(steps), _f64(grid)
    thr = nearest_thresholds(g)
    bad = (ctypes.c_uint64 * 2)(0, 0)
    check(lib.pb200_selftest_exact(
        ctypes.c_int(device), ctypes.c_int64(int(n)), ctypes.c_uint64(int(seed)), _dp(s),
        ctypes.c_int(len(s)), _dp(g), _dp(thr), ctypes.c_int(len(g)), bad))
    return int(bad[0]), int(bad[1])
