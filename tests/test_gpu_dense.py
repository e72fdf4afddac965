"""GPU parity of the dense-convolution accumulate path (csrc/dense_kernels.cu): isotopes whose
co-add groups fill a large share of the fine grid are evaluated as a dense strided convolution
instead of per-group gathers.  Forced here on small grids with PB200_DENSE_MIN_OCC; checked
against the CPU oracle (1e-10 of each layer's peak, identical nadd/nskip/neval counters) and
against the gather path of the same engine (rounding only)."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

TOL_PEAK = 1e-10


def _engine(case):
    from pyratbay_b200.engine import Engine
    eng = Engine(0)
    eng.set_grid(case.spec.wn, case.spec.own, case.spec.odivisors)
    eng.set_species(case.atm.mol_radius, case.atm.mol_mass, case.iso_atm_index,
                    case.iso_mass, case.iso_ratio)
    eng.set_lines(case.lwn, case.elow, case.gf, case.isoid)
    eng.set_voigt(case.lorentz, case.doppler, case.size, case.index, case.profile, case.cutoff)
    return eng


def _oracle(case, temps, dens, isoz, add, iso_iext=None, nextinct=1):
    orc = helpers.oracle_module()
    nrows = 1 if add else nextinct
    out = np.zeros((len(temps), nrows, case.spec.nwave))
    cnt = np.zeros((len(temps), 4), np.int64)
    for u in range(len(temps)):
        ext = np.zeros((nextinct, case.spec.nwave))
        orc.extinction(ext, *case.unit_args(temps[u], dens[u], isoz[u], iso_iext), 0, int(add), 0,
                       counters=cnt[u])
        out[u] = ext[:nrows]
    return out, cnt


def _peak_err(got, want):
    peak = np.max(np.abs(want), axis=-1, keepdims=True)
    peak[peak == 0] = 1.0
    return np.max(np.abs(got - want) / peak)


CASES = [
    # S = 360 (11.25 lane blocks), cutoff-limited windows, ofactor 1 .. 360 over the pressures
    dict(nlines=100_000, wnlow=9000.0, wnhigh=9400.0, wnstep=0.25, wnosamp=360, cutoff=10.07,
         extent=60.0, nlayers=11),
    # many skipped lines
    dict(nlines=100_000, wnlow=9000.0, wnhigh=9400.0, wnstep=0.25, wnosamp=360, cutoff=10.07,
         extent=60.0, nlayers=7, ethresh=1e-6),
    # no fixed cutoff: windows set by the profile sizes
    dict(nlines=60_000, wnlow=9000.0, wnhigh=9400.0, wnstep=0.25, wnosamp=240, cutoff=0.0,
         extent=8.0, nlayers=9),
    # heavy co-adding (4 lines per fine cell), S = 120
    dict(nlines=800_000, wnlow=9000.0, wnhigh=9400.0, wnstep=0.25, wnosamp=120, cutoff=12.13,
         extent=40.0, nlayers=7),
    # the table benchmark's sampling: 0.33 cm-1 output step, S = 840, cutoff 25 cm-1
    dict(nlines=200_000, wnlow=8000.0, wnhigh=8600.0, wnstep=0.33, wnosamp=840, cutoff=25.0,
         extent=300.0, nlayers=9),
]
# (cutoff/(ownstep*ofactor) must not be within 1e-6 of an integer for the dense path: the
# reference's (int)(idwn +- cutoff/dwnstep) is then position dependent and the engine keeps
# such units on the gather path; test_near_integer_cutoff_steps_stay_on_the_gather_path)


@pytest.mark.parametrize("form", ["ws", "2cta"])
@pytest.mark.parametrize("kwargs", CASES)
def test_dense_path_matches_oracle_and_gather(kwargs, form, monkeypatch):
    # both forms of the dense kernel: warp-specialised (producer warps, double-buffered tiles)
    # and two CTAs per SM
    monkeypatch.setenv("PB200_DENSE_WS", "1" if form == "ws" else "0")
    case = helpers.synthetic_case(**kwargs)
    temps, dens = case.atm.temp, case.atm.d
    isoz = helpers.partition(case, temps).T
    monkeypatch.setenv("PB200_DENSE_MIN_OCC", "0.0005")      # every isotope with lines
    monkeypatch.setenv("PB200_DENSE_MIN_SPAN", "1")          # ... and all of their cells
    for add in (0, 1):
        args = (temps, dens, isoz, case.iso_mol_index, 1, case.ethresh, add, 0)
        want, wcnt = _oracle(case, temps, dens, isoz, add)
        monkeypatch.setenv("PB200_DENSE", "1")
        eng = _engine(case)
        got, cnt = eng.extinction_batch(*args, counters=True)
        used = eng.dense_units()
        eng.close()
        assert used > 0, "the dense path was not taken"
        assert np.array_equal(cnt[:, :4], wcnt)
        assert _peak_err(got, want) < TOL_PEAK
        monkeypatch.setenv("PB200_DENSE", "0")
        eng = _engine(case)
        gather = eng.extinction_batch(*args)
        assert eng.dense_units() == 0
        eng.close()
        assert _peak_err(got, gather) < 1e-13


def test_dense_main_isotope_gather_for_the_rest(monkeypatch):
    """Threshold between the isotopes' occupancies: 75 % of the lines go dense, the minor
    isotopes through the gather kernels, both into the same output rows; per-species rows."""
    case = helpers.synthetic_case(nlines=150_000, wnlow=9000.0, wnhigh=9400.0, wnstep=0.25,
                                  wnosamp=360, cutoff=10.07, extent=60.0, nlayers=7)
    temps, dens = case.atm.temp, case.atm.d
    isoz = helpers.partition(case, temps).T
    nfine = len(case.spec.own)
    occ = np.bincount(case.isoid, minlength=4) / nfine
    monkeypatch.setenv("PB200_DENSE_MIN_OCC", f"{0.5 * (occ[0] + occ[1]) * 0.8:.5f}")
    eng = _engine(case)
    # two output rows: isotopes 0, 2 -> row 0; isotope 1 -> row 1; isotope 3 skipped
    iext = np.array([0, 1, 0, -1])
    got = eng.extinction_batch(temps, dens, isoz, iext, 2, case.ethresh, 0, 0)
    used = eng.dense_units()
    eng.close()
    # the main isotope on its own plane, isotope 2 (same output row) merged into it where it
    # selects the same profile: at most two isotopes per unit on the dense path
    assert 0 < used <= 2 * len(temps)
    want, _ = _oracle(case, temps, dens, isoz, 0, iso_iext=iext, nextinct=2)
    assert got.shape == want.shape
    assert _peak_err(got, want) < TOL_PEAK


def test_dense_path_lines_on_grid_points(monkeypatch):
    """Lines exactly on fine-grid samples and exactly half way between them: the cases where
    the dynamic index of a line sits on a rounding edge (anomalous cells)."""
    base = helpers.synthetic_case(nlines=50_000, wnlow=9000.0, wnhigh=9400.0, wnstep=0.25,
                                  wnosamp=360, cutoff=10.07, extent=60.0, nlayers=9)
    own = base.spec.own
    rng = np.random.default_rng(5)
    lwn = base.lwn.copy()
    iso = base.isoid
    # snap 60 % of the lines: a third onto samples, a third onto midpoints, a third just below
    pick = rng.random(len(lwn)) < 0.6
    idx = np.clip(np.searchsorted(own, lwn), 1, len(own) - 2)
    kind = rng.integers(0, 3, len(lwn))
    snapped = np.where(kind == 0, own[idx], np.where(kind == 1, 0.5 * (own[idx] + own[idx - 1]),
                                                     np.nextafter(own[idx], 0)))
    lwn = np.where(pick, snapped, lwn)
    order = np.lexsort((lwn, iso))
    case = helpers.Case(**{**base.__dict__, "lwn": lwn[order], "elow": base.elow[order],
                           "gf": base.gf[order], "isoid": iso[order]})
    temps, dens = case.atm.temp, case.atm.d
    isoz = helpers.partition(case, temps).T
    monkeypatch.setenv("PB200_DENSE_MIN_OCC", "0.0005")
    eng = _engine(case)
    got, cnt = eng.extinction_batch(temps, dens, isoz, case.iso_mol_index, 1, case.ethresh, 0, 0,
                                    counters=True)
    assert eng.dense_units() > 0
    eng.close()
    want, wcnt = _oracle(case, temps, dens, isoz, 0)
    assert np.array_equal(cnt[:, :4], wcnt)
    assert _peak_err(got, want) < TOL_PEAK


def test_near_integer_cutoff_steps_stay_on_the_gather_path(monkeypatch):
    """cutoff/dwnstep within rounding of an integer (the reference's defaults: 25 cm-1, 1 cm-1
    steps): (int)(idwn +- cutoff/dwnstep) depends on the magnitude of idwn, the windows are not
    translation invariant and the engine must not take the dense path."""
    case = helpers.synthetic_case(nlines=100_000, wnlow=9000.0, wnhigh=9400.0, wnstep=0.25,
                                  wnosamp=360, cutoff=10.0, extent=60.0, nlayers=5)
    temps, dens = case.atm.temp, case.atm.d
    isoz = helpers.partition(case, temps).T
    monkeypatch.setenv("PB200_DENSE_MIN_OCC", "0.0005")
    eng = _engine(case)
    got = eng.extinction_batch(temps, dens, isoz, case.iso_mol_index, 1, case.ethresh, 0, 0)
    assert eng.dense_units() == 0
    eng.close()
    want, _ = _oracle(case, temps, dens, isoz, 0)
    assert _peak_err(got, want) < TOL_PEAK


@pytest.mark.parametrize("min_span", ["1", "24"])
@pytest.mark.parametrize("merge", ["1", "0"])
def test_minor_isotopes_merged_into_the_main_plane(min_span, merge, monkeypatch):
    """One dense plane for the main isotope plus the minor isotopes of the same species where
    they select the same Voigt profile; the cells where their Doppler sample differs from the
    main isotope's, and narrow footprints, stay with the gather kernels.  Same results as the
    oracle with identical counters, with and without merging."""
    case = helpers.synthetic_case(nlines=150_000, wnlow=9000.0, wnhigh=9400.0, wnstep=0.25,
                                  wnosamp=360, cutoff=10.07, extent=60.0, nlayers=9, ndop=40)
    temps, dens = case.atm.temp, case.atm.d
    isoz = helpers.partition(case, temps).T
    nfine = len(case.spec.own)
    occ = np.bincount(case.isoid, minlength=4) / nfine
    monkeypatch.setenv("PB200_DENSE_MIN_OCC", f"{0.5 * (occ[0] + occ[1]) * 0.8:.5f}")
    monkeypatch.setenv("PB200_DENSE_MIN_SPAN", min_span)
    monkeypatch.setenv("PB200_DENSE_MERGE", merge)
    for add in (0, 1):
        eng = _engine(case)
        got, cnt = eng.extinction_batch(temps, dens, isoz, case.iso_mol_index, 1, case.ethresh,
                                        add, 0, counters=True)
        used = eng.dense_units()
        eng.close()
        want, wcnt = _oracle(case, temps, dens, isoz, add)
        assert np.array_equal(cnt[:, :4], wcnt)
        assert _peak_err(got, want) < TOL_PEAK
        if merge == "1":
            assert used > len(temps)          # more than the main isotope took the dense path
        else:
            assert 0 < used <= len(temps)


def test_two_dense_planes_and_merged_minors(monkeypatch):
    """Occupancy threshold between the second and third isotope: two isotopes get a dense plane
    of their own, the two rarest are merged into the main one's."""
    case = helpers.synthetic_case(nlines=150_000, wnlow=9000.0, wnhigh=9400.0, wnstep=0.25,
                                  wnosamp=360, cutoff=10.07, extent=60.0, nlayers=7, ndop=40)
    temps, dens = case.atm.temp, case.atm.d
    isoz = helpers.partition(case, temps).T
    occ = np.bincount(case.isoid, minlength=4) / len(case.spec.own)
    monkeypatch.setenv("PB200_DENSE_MIN_OCC", f"{0.5 * (occ[1] + occ[2]) * 0.8:.5f}")
    monkeypatch.setenv("PB200_DENSE_MIN_SPAN", "8")
    eng = _engine(case)
    got, cnt = eng.extinction_batch(temps, dens, isoz, case.iso_mol_index, 1, case.ethresh, 0, 0,
                                    counters=True)
    used = eng.dense_units()
    eng.close()
    want, wcnt = _oracle(case, temps, dens, isoz, 0)
    assert np.array_equal(cnt[:, :4], wcnt)
    assert _peak_err(got, want) < TOL_PEAK
    assert used > 2 * len(temps)        # two planes plus merged isotopes


def test_dense_path_with_chunked_strength_passes(monkeypatch):
    """Strength passes forced into chunks of 2 temperatures and gather launches of at most 3
    units: the dense path's per-pass state (Doppler segments of main and minor isotopes, dense
    planes) is chunk-relative; same bits as the unchunked batch."""
    case = helpers.synthetic_case(nlines=150_000, wnlow=9000.0, wnhigh=9400.0, wnstep=0.25,
                                  wnosamp=360, cutoff=10.07, extent=60.0, nlayers=7, ndop=40)
    temps, dens = case.atm.temp, case.atm.d
    isoz = helpers.partition(case, temps).T
    occ = np.bincount(case.isoid, minlength=4) / len(case.spec.own)
    monkeypatch.setenv("PB200_DENSE_MIN_OCC", f"{0.5 * (occ[0] + occ[1]) * 0.8:.5f}")
    args = (temps, dens, isoz, case.iso_mol_index, 1, case.ethresh, 0, 0)
    eng = _engine(case)
    base = eng.extinction_batch(*args)
    assert eng.dense_units() > len(temps)
    eng.close()
    monkeypatch.setenv("PB200_TP_CHUNK", "2")
    monkeypatch.setenv("PB200_MAX_UNITS_PER_LAUNCH", "3")
    eng = _engine(case)
    chunked = eng.extinction_batch(*args)
    assert eng.dense_units() > len(temps)
    eng.close()
    assert np.array_equal(chunked, base)
    want, _ = _oracle(case, temps, dens, isoz, 0)
    assert _peak_err(base, want) < TOL_PEAK
