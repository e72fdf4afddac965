"""GPU parity at the sizes of BASELINE.json configs[2] (1e7- and 1e8-line cross-section tables,
constant-step and constant-R grids) and of the run_batch branches that only large batches
reach (strength-pass chunking, launch splits, > 65535 units, 32-bit offset fallback).

A whole table row costs the CPU oracle ~100 s at 1e8 lines, so rows are spot-checked on
spectral sub-windows: the oracle runs on the FULL grids (identical own0, ofactor, fine indices)
with the lines that can reach the window (window +- cutoff +- margin) and the output samples
of the window are compared.  ethresh is the default 1e-30, for which the per-row maximum does
not skip anything in either line set (the skip counters are asserted to be zero on both
sides); identical line selection is checked with a second engine holding the same line subset
(nadd/nskip/neval/sample counters equal to the oracle's).
"""
import numpy as np
import pytest

import helpers
from pyratbay_b200 import constants as pc

pytestmark = pytest.mark.gpu

TOL_PEAK = 1e-10     # relative to the peak extinction (here: of the compared window)


def _table_engine(nlines, resolution=None, nwave=100_000):
    """Engine + static arrays of the bench table workload (workloads.table_workload)."""
    from pyratbay_b200 import workloads
    from pyratbay_b200.engine import Engine
    from pyratbay_b200.spectrum import Spectrum
    from pyratbay_b200.voigt import Voigt
    w = workloads.table_workload(nlines, nwave=nwave)
    if resolution:
        # constant-R variant of configs[2]: reference defaults wnstep 1.0, wnosamp by the 4e-4 rule
        spec = Spectrum(wnlow=w.inputs["wnlow"], wnhigh=w.inputs["wnhigh"], wnstep=1.0,
                        resolution=resolution)
    else:
        spec = Spectrum(wnlow=w.inputs["wnlow"], wnhigh=w.inputs["wnhigh"], wnstep=w.wnstep,
                        wnosamp=w.wnosamp)
    lwn, elow, gf, iso, _ = w.make_lines()
    eng = Engine(0)
    eng.set_grid(spec.wn, spec.own, spec.odivisors)
    eng.set_species(w.atm.mol_radius, w.atm.mol_mass, w.iso_atm_index, w.db.iso_mass,
                    w.db.iso_ratio)
    eng.set_lines(lwn, elow, gf, iso.astype(np.int64))
    voigt = Voigt(spec, w.atm, w.iso_atm_index, eng, tmin=w.inputs["tmin"],
                  tmax=w.inputs["tmax"])
    return w, spec, eng, voigt, (lwn, elow, gf, iso.astype(np.int64))


def _units(w, pairs):
    from pyratbay_b200 import workloads
    itemp = np.array([p[0] for p in pairs])
    ilayer = np.array([p[1] for p in pairs])
    temps = w.temps[itemp]
    atm = w.atm
    dens = atm.vmr[ilayer] * atm.press[ilayer, None] * pc.bar / (pc.k * temps[:, None])
    isoz = workloads.partition(w.db, temps)
    return temps, dens, isoz


def _spot_check(w, spec, eng, voigt, lines, pairs, windows, resolution):
    """Rows of `pairs` = [(itemp, ilayer)] from the full-list engine vs the oracle on the
    sub-windows `windows` = [(first output, count)]."""
    from pyratbay_b200.engine import Engine
    orc = helpers.oracle_module()
    lwn, elow, gf, iso = lines
    temps, dens, isoz = _units(w, pairs)
    interp = 1 if resolution else 0
    got = eng.extinction_batch(temps, dens, isoz, w.iso_mol_index, 1, 1e-30, 0, interp)

    # lines that can reach a window: centre within cutoff (+ 2 cm-1 margin) of its ends
    reach = voigt.cutoff + 2.0
    keep = np.zeros(len(lwn), bool)
    for first, count in windows:
        keep |= (lwn >= spec.wn[first] - reach) & (lwn <= spec.wn[first + count - 1] + reach)
    sub = (lwn[keep], elow[keep], gf[keep], iso[keep])

    profile = voigt.profile                     # host copy of the device-built table
    sub_eng = Engine(0)
    sub_eng.set_grid(spec.wn, spec.own, spec.odivisors)
    sub_eng.set_species(w.atm.mol_radius, w.atm.mol_mass, w.iso_atm_index, w.db.iso_mass,
                        w.db.iso_ratio)
    sub_eng.set_lines(*sub)
    sub_eng.set_voigt(voigt.lorentz, voigt.doppler, voigt.size, voigt.index, profile,
                      voigt.cutoff)
    sub_got, sub_cnt = sub_eng.extinction_batch(temps, dens, isoz, w.iso_mol_index, 1, 1e-30, 0,
                                                interp, counters=True)
    sub_eng.close()

    worst = 0.0
    for u in range(len(pairs)):
        ext = np.zeros((1, spec.nwave))
        cnt = np.zeros(4, np.int64)
        orc.extinction(ext, profile, voigt.size, voigt.index, voigt.lorentz, voigt.doppler,
                       spec.wn, spec.own, spec.odivisors, dens[u], w.atm.mol_radius,
                       w.atm.mol_mass, w.iso_atm_index, w.db.iso_mass, w.db.iso_ratio, isoz[u],
                       w.iso_mol_index, *sub, voigt.cutoff, 1e-30, temps[u], 0, 0, interp,
                       counters=cnt)
        assert np.array_equal(sub_cnt[u, :4], cnt), "nadd/nskip/neval/sample counters differ"
        assert cnt[1] == 0                      # nothing skipped: the subset's kmax is immaterial
        # the subset engine reproduces the oracle over the WHOLE grid ...
        peak = ext[0].max()
        assert peak > 0
        assert np.max(np.abs(sub_got[u, 0] - ext[0])) / peak < TOL_PEAK
        # ... and the full-list row equals it wherever every contributing line is in the subset
        for first, count in windows:
            sl = slice(first, first + count)
            wpeak = ext[0, sl].max()
            assert wpeak > 0
            err = np.max(np.abs(got[u, 0, sl] - ext[0, sl])) / wpeak
            worst = max(worst, err)
            assert err < TOL_PEAK, (pairs[u], first, err)
    return worst


PAIRS = [(0, 0), (3, 45), (5, 20), (10, 35), (15, 10), (19, 50)]   # (itemp, ilayer)


@pytest.mark.parametrize("nlines", [10_000_000, 100_000_000])
def test_table_rows_match_oracle_at_scale(nlines):
    """Constant-step grid of the bench table (wide <3,16> instantiation of the chunk kernel,
    slot-major multi-pass path): six (T,p) rows x three spectral windows."""
    w, spec, eng, voigt, lines = _table_engine(nlines)
    assert eng.line_stats()["in_window"] > 0.99 * nlines
    windows = [(40, 120), (50_000, 120), (99_700, 120)]
    worst = _spot_check(w, spec, eng, voigt, lines, PAIRS, windows, None)
    print(f"table {nlines:.0e} lines: max |d|/window peak = {worst:.2e}")
    eng.close()


def test_constant_resolution_table_rows_match_oracle_at_scale():
    """configs[2], constant-R variant (R = 20000 over 0.3-30 um, 2-point interpolation)."""
    w, spec, eng, voigt, lines = _table_engine(10_000_000, resolution=20000.0)
    assert spec.interpolate and spec.nwave > 90_000
    windows = [(100, 150), (spec.nwave // 2, 150), (spec.nwave - 400, 150)]
    worst = _spot_check(w, spec, eng, voigt, lines, PAIRS, windows, 20000.0)
    print(f"constant-R table: max |d|/window peak = {worst:.2e}")
    eng.close()


# ------------------------------------------------------------------- run_batch branches
def _small_batch(n_units_t=10, reps=4):
    """40 units with 10 distinct temperatures on a small synthetic case."""
    case = helpers.synthetic_case(nlines=30000, nlayers=n_units_t)
    temps = np.tile(case.atm.temp, reps)
    dens = np.tile(case.atm.d, (reps, 1)) * np.repeat(10.0 ** np.arange(reps), n_units_t)[:, None]
    isoz = np.tile(helpers.partition(case, case.atm.temp).T, (reps, 1))
    return case, temps, dens, isoz


def _engine(case):
    from pyratbay_b200.engine import Engine
    eng = Engine(0)
    eng.set_grid(case.spec.wn, case.spec.own, case.spec.odivisors)
    eng.set_species(case.atm.mol_radius, case.atm.mol_mass, case.iso_atm_index,
                    case.iso_mass, case.iso_ratio)
    eng.set_lines(case.lwn, case.elow, case.gf, case.isoid)
    eng.set_voigt(case.lorentz, case.doppler, case.size, case.index, case.profile, case.cutoff)
    return eng


def test_strength_pass_chunking_and_launch_splits_are_bit_identical(monkeypatch):
    """run_batch with the strength passes forced into chunks (PB200_TP_CHUNK: the per-chunk
    rewrite of the staging block) and with the accumulate launches split
    (PB200_MAX_UNITS_PER_LAUNCH) gives the bits of the single-launch batch; ksplit (several
    CTAs per tile, partial sums) agrees to rounding."""
    case, temps, dens, isoz = _small_batch()
    args = (temps, dens, isoz, case.iso_mol_index, 1, case.ethresh, 0, 0)

    def run():
        eng = _engine(case)
        out = eng.extinction_batch(*args)
        eng.close()
        return out
    base = run()
    monkeypatch.setenv("PB200_TP_CHUNK", "3")
    assert np.array_equal(run(), base)
    monkeypatch.delenv("PB200_TP_CHUNK")
    monkeypatch.setenv("PB200_MAX_UNITS_PER_LAUNCH", "7")
    assert np.array_equal(run(), base)
    monkeypatch.setenv("PB200_TP_CHUNK", "4")
    assert np.array_equal(run(), base)
    monkeypatch.delenv("PB200_TP_CHUNK")
    monkeypatch.delenv("PB200_MAX_UNITS_PER_LAUNCH")
    peak = base.max(axis=-1, keepdims=True)
    for ks in ("1", "4"):
        monkeypatch.setenv("PB200_KSPLIT", ks)
        monkeypatch.setenv("PB200_MAX_UNITS_PER_LAUNCH", "9")   # partial buffer reused per launch
        assert np.max(np.abs(run() - base) / peak) < 1e-13
    # against the oracle, first repetition
    orc = helpers.oracle_module()
    for u in (0, 5, 9):
        ext = np.zeros((1, case.spec.nwave))
        orc.extinction(ext, *case.unit_args(temps[u], dens[u], isoz[u]), 0, 0, 0)
        assert np.max(np.abs(base[u, 0] - ext[0])) / ext[0].max() < TOL_PEAK


def test_more_than_65535_units_in_one_batch():
    """A batch larger than one launch's grid.y limit (65535 units) on a tiny grid."""
    case = helpers.synthetic_case(nlines=300, wnlow=9000.0, wnhigh=9040.0, nlayers=9)
    n = 70_000
    reps = -(-n // 9)
    temps = np.tile(case.atm.temp, reps)[:n]
    dens = np.tile(case.atm.d, (reps, 1))[:n]
    isoz = np.tile(helpers.partition(case, case.atm.temp).T, (reps, 1))[:n]
    eng = _engine(case)
    out = eng.extinction_batch(temps, dens, isoz, case.iso_mol_index, 1, case.ethresh, 1, 0)
    eng.close()
    assert out.shape == (n, 1, case.spec.nwave)
    first = out[:9]
    assert np.array_equal(out[:9 * (n // 9)].reshape(n // 9, 9, -1), np.tile(first[:, 0], (n // 9, 1, 1)))
    assert np.array_equal(out[9 * (n // 9):, 0], first[:n % 9, 0])
    orc = helpers.oracle_module()
    for u in range(9):
        ext = np.zeros((1, case.spec.nwave))
        orc.extinction(ext, *case.unit_args(temps[u], dens[u], isoz[u]), 0, 1, 0)
        assert np.max(np.abs(out[u, 0] - ext[0])) / ext[0].max() < TOL_PEAK


def test_large_voigt_table_falls_back_to_64bit_offsets(monkeypatch):
    """An output-stride table too long for the packed 32-bit offsets (> 34 GB; forced here with
    PB200_PACK_LIMIT) must take the strided gather with 64-bit addresses, not wrap around."""
    case = helpers.synthetic_case()
    temps, dens = case.atm.temp, case.atm.d
    isoz = helpers.partition(case, temps).T
    args = (temps, dens, isoz, case.iso_mol_index, 1, case.ethresh, 0, 0)
    monkeypatch.setenv("PB200_ACC_MODE", "strided")
    eng = _engine(case)
    strided = eng.extinction_batch(*args)
    eng.close()
    monkeypatch.delenv("PB200_ACC_MODE")
    monkeypatch.setenv("PB200_PACK_LIMIT", "1000")
    eng = _engine(case)
    forced = eng.extinction_batch(*args)
    eng.close()
    assert np.array_equal(forced, strided)
    want = np.zeros_like(forced)
    orc = helpers.oracle_module()
    for u in range(len(temps)):
        ext = np.zeros((1, case.spec.nwave))
        orc.extinction(ext, *case.unit_args(temps[u], dens[u], isoz[u]), 0, 0, 0)
        want[u] = ext
    peak = want.max(axis=-1, keepdims=True)
    assert np.max(np.abs(forced - want) / peak) < TOL_PEAK


# ------------------------------------------------------------------- compute_opacity, device
def test_compute_opacity_leaves_the_table_on_the_device(tmp_path):
    """extinction.compute_opacity: rows written straight into the device table (no host
    bounce), host copy and .npz equal to it; rows equal a direct batch call."""
    import torch
    from pyratbay_b200 import io, tli as ptli, workloads
    from pyratbay_b200.pyrat import Pyrat
    w = workloads.table_workload(200_000, ntemp=4, nlayers=7, nwave=2000, wl_low_um=1.0,
                                 wl_high_um=1.2)
    path = str(tmp_path / "lines.tli")
    wn, elow, gf, iso, counts = w.make_lines()
    ptli.write_tli(path, [w.db], [{"wn": wn, "elow": elow, "gf": gf, "iso_id": iso,
                                   "n_lines_iso": counts}], w.inputs["wnlow"], w.inputs["wnhigh"])
    cs = str(tmp_path / "table.npz")
    pyrat = Pyrat(dict(w.inputs, tlifile=[path], sampled_cs=[cs]), atm=w.atm, device=0)
    pyrat.compute_opacity()
    ex = pyrat.ex
    assert isinstance(ex.etable_dev, torch.Tensor) and ex.etable_dev.is_cuda
    assert ex.etable.shape == (4, 7, 2000)
    assert np.array_equal(ex.etable_dev.cpu().numpy(), ex.etable)
    assert np.array_equal(io.read_opacity(cs, extract="opacity"), ex.etable)
    temps, dens, isoz = _units(w, [(t, p) for t in range(4) for p in range(7)])
    direct = pyrat.engine.extinction_batch(temps, dens, isoz, w.iso_mol_index, 1, 1e-30, 0, 0)
    assert np.array_equal(direct[:, 0].reshape(4, 7, 2000), ex.etable)
    assert ex.etable.max() > 0


@pytest.mark.parametrize("factor", ["0", "0.1", "0.001"])
def test_constant_resolution_dynamic_grid_path(factor, monkeypatch):
    """Constant-R grid: units evaluated on their dynamic grid with the chunk kernel and then
    interpolated (PB200_DYN_FACTOR: never / default split / every unit that qualifies) against the
    oracle, with equal counters; add = 0 and 1."""
    monkeypatch.setenv("PB200_DYN_FACTOR", factor)
    case = helpers.synthetic_case(nlines=60000, wnlow=9000.0, wnhigh=9400.0, resolution=9000.0,
                                  nlayers=11)
    temps, dens = case.atm.temp, case.atm.d
    isoz = helpers.partition(case, temps).T
    orc = helpers.oracle_module()
    for add in (0, 1):
        eng = _engine(case)
        got, cnt = eng.extinction_batch(temps, dens, isoz, case.iso_mol_index, 1, case.ethresh,
                                        add, 1, counters=True)
        eng.close()
        for u in range(len(temps)):
            ext = np.zeros((1, case.spec.nwave))
            c = np.zeros(4, np.int64)
            orc.extinction(ext, *case.unit_args(temps[u], dens[u], isoz[u]), 0, add, 1, counters=c)
            assert np.array_equal(cnt[u, :4], c)
            assert np.max(np.abs(got[u, 0] - ext[0])) / ext[0].max() < TOL_PEAK
