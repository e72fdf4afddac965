"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle
(oracle/lbl_oracle.c) and against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).

Tolerances (BASELINE.json north_star): extinction within 1e-10 of each layer's peak in fp64,
identical ethresh line selection (checked through the nskip/neval counters); Voigt table
within 1e-12 relative; interp_ec bit-exact.
"""
import os

import numpy as np
import pytest

import helpers
from pyratbay_b200 import constants as pc

pytestmark = pytest.mark.gpu

TOL_PEAK = 1e-10     # relative to each layer's peak extinction
TOL_VOIGT = 1e-12    # relative, per profile sample


def _engine_for(case, profile="device"):
    """Engine loaded with the static inputs of `case`; Voigt table built on the device
    ('device') or adopted from the oracle's host table ('host')."""
    from pyratbay_b200.engine import Engine
    eng = Engine(0)
    eng.set_grid(case.spec.wn, case.spec.own, case.spec.odivisors)
    eng.set_species(case.atm.mol_radius, case.atm.mol_mass, case.iso_atm_index,
                    case.iso_mass, case.iso_ratio)
    eng.set_lines(case.lwn, case.elow, case.gf, case.isoid)
    if profile == "device":
        size = case.size_in.copy()
        index = np.zeros_like(size)
        eng.build_voigt(case.lorentz, case.doppler, case.spec.ownstep, size, index,
                        case.cutoff)
        assert np.array_equal(size, case.size)
        assert np.array_equal(index, case.index)
    else:
        eng.set_voigt(case.lorentz, case.doppler, case.size, case.index, case.profile,
                      case.cutoff)
    return eng


def _oracle_units(case, temps, dens, isoz, iso_iext, nextinct, add, resolution):
    orc = helpers.oracle_module()
    nrows = 1 if add else nextinct
    out = np.zeros((len(temps), nrows, case.spec.nwave))
    cnt = np.zeros((len(temps), 4), np.int64)
    for u in range(len(temps)):
        ext = np.zeros((nextinct, case.spec.nwave))
        orc.extinction(ext, *case.unit_args(temps[u], dens[u], isoz[u], iso_iext),
                       0, int(add), int(resolution), counters=cnt[u])
        out[u] = ext[:nrows]
    return out, cnt


def _peak_err(got, want):
    peak = np.max(np.abs(want), axis=-1, keepdims=True)
    peak[peak == 0] = 1.0
    return np.max(np.abs(got - want) / peak)


# ---------------------------------------------------------------------------- Voigt grid
def test_voigt_grid_matches_oracle_and_reference_golden():
    from pyratbay_b200.engine import voigt_grid
    case = helpers.mock_case()
    g = helpers.golden("mock_voigt.npz")
    size = case.size_in.copy()
    index = np.zeros_like(size)
    profile = np.zeros(case.voigt.profile_len)
    assert voigt_grid(profile, size, index, case.lorentz, case.doppler,
                      case.spec.ownstep) == 1
    assert np.array_equal(size, g["size"])
    assert np.array_equal(index, g["index"])
    # against the oracle, every sample
    ref = case.profile
    nz = ref > 0
    assert np.max(np.abs(profile[nz] / ref[nz] - 1.0)) < TOL_VOIGT
    assert np.all(profile[~nz] == 0.0)
    # against the reference's own table (strided golden samples, whole edge profiles)
    stride = int(g["stride"])
    want = g["profile_strided"]
    got = profile[::stride]
    nz = want > 0
    assert np.max(np.abs(got[nz] / want[nz] - 1.0)) < TOL_VOIGT
    for key, (m, n) in {"profile_first": (0, 0), "profile_mid": (40, 10),
                        "profile_last": (-1, -1)}.items():
        seg = profile[index[m, n]:index[m, n] + 2 * size[m, n] + 1]
        assert np.max(np.abs(seg / g[key] - 1.0)) < TOL_VOIGT
    assert abs(np.sum(profile) / float(g["profile_sum"]) - 1.0) < 1e-12


def test_voigt_known_answer_of_reference_test_str():
    """The 1.1-1.7 um H2O Voigt grid pinned by the reference's tests/test_str.py:338-366."""
    from pyratbay_b200.spectrum import Spectrum
    from pyratbay_b200.voigt import Voigt
    from pyratbay_b200.engine import Engine
    g = helpers.golden("voigt_h2o_1.1-1.7um.npz")
    spec = Spectrum(wl_low=1.1 * pc.um, wl_high=1.7 * pc.um, wnstep=1.0, wnosamp=2160)
    atm = helpers.mock_atmosphere()
    eng = Engine(0)
    iso_atm_index = np.full(4, list(atm.species).index("H2O"))
    # no tmin/tmax: the Voigt width bounds default to 100-3000 K (pyrat/voigt.py:35-36)
    v = Voigt(spec, atm, iso_atm_index, eng, extent=100.0, cutoff=25.0)
    # numbers printed in test_str.py:352-365
    assert list(v.size[0, [0, 1, -2, -1]]) == [1072, 1120, 8687, 9074]
    assert list(v.size[-1, [0, -1]]) == [54000, 54000]
    assert list(v.index[0, [0, 1, -2, -1]]) == [0, 2145, 341896, 359271]
    assert list(v.index[1, [0, 1, -2, -1]]) == [377420, 379565, 719318, 736693]
    assert v.index[-2, 0] == 46933096 and v.index[-1, -1] == 47041097
    assert np.array_equal(v.size, g["size"]) and np.array_equal(v.index, g["index"])
    prof = v.profile
    assert len(prof) == int(g["profile_len"])
    first = prof[v.index[0, 0]:v.index[0, 0] + 2 * v.size[0, 0] + 1]
    last = prof[v.index[-1, -1]:v.index[-1, -1] + 2 * v.size[-1, -1] + 1]
    assert ["%.5e" % x for x in first[[0, 1, -2, -1]]] == \
        ["2.85914e-08", "2.86448e-08", "2.86448e-08", "2.85914e-08"]
    assert ["%.5e" % x for x in last[[0, 1, -2, -1]]] == \
        ["4.99389e-03", "4.99404e-03", "4.99404e-03", "4.99389e-03"]
    want = g["profile_strided"]
    got = prof[::int(g["stride"])]
    nz = want > 0
    assert np.max(np.abs(got[nz] / want[nz] - 1.0)) < TOL_VOIGT
    assert np.all(got[~nz] == 0.0)
    # the whole known-answer string of the reference's test
    assert str(v) == str(g["voigt_str"])
    eng.close()


# ------------------------------------------------------------------- reference golden tables
@pytest.mark.parametrize("resolution,golden_file", [
    (None, "mock_opacity_table.npz"), (15000.0, "mock_opacity_table_R.npz")])
def test_compute_opacity_matches_reference_table(tmp_path, resolution, golden_file):
    """configs[0]: pbay -c opacity on the mock HITRAN H2O TLI, through the Pyrat-shaped API."""
    import pyratbay_b200 as pb
    g = helpers.golden(golden_file)
    inputs = dict(
        runmode="opacity", tlifile=os.path.join(helpers.GOLDEN, "mock_hitran_h2o.tli"),
        wl_low=1.00 * pc.um, wl_high=1.01 * pc.um, wnstep=1.0, wnosamp=2160,
        resolution=resolution, voigt_extent=100.0, tmin=300.0, tmax=3000.0, tstep=300.0,
        sampled_cs=[str(tmp_path / "table.npz")], verb=0)
    pyrat = pb.Pyrat(inputs, atm=helpers.mock_atmosphere())
    pyrat.compute_opacity()
    units, species, temp, press, wn, table = pb.io.read_opacity(inputs["sampled_cs"][0])
    assert species == "H2O" and units["pressure"] == "bar"
    assert np.array_equal(temp, g["temp"]) and np.array_equal(wn, g["wn"])
    assert np.array_equal(press, g["press"])
    assert table.shape == g["etable"].shape
    assert _peak_err(table, g["etable"]) < TOL_PEAK


def test_forward_model_matches_reference():
    """LBL branch of pyrat.run: co-added extinction for all layers, one-layer per-species
    extinction and skip_mol (golden from the reference's Line_By_Line)."""
    import pyratbay_b200 as pb
    g = helpers.golden("mock_forward.npz")
    inputs = dict(
        tlifile=os.path.join(helpers.GOLDEN, "mock_hitran_h2o.tli"),
        wl_low=1.00 * pc.um, wl_high=1.01 * pc.um, wnstep=1.0, wnosamp=2160,
        voigt_extent=100.0, tmin=300.0, tmax=3000.0, tstep=300.0, verb=0)
    pyrat = pb.Pyrat(inputs, atm=helpers.mock_atmosphere())
    assert np.array_equal(pyrat.atm.d, g["d"])
    ec = pyrat.calc_lbl_extinction()
    assert _peak_err(ec, g["ec_all"]) < TOL_PEAK
    ec31, label = pyrat.get_ec(31)
    assert label == ["H2O"]
    assert _peak_err(ec31, g["ec_layer31"]) < TOL_PEAK
    skipped = pyrat.calc_lbl_extinction(skip_mol=["H2O"])
    assert np.all(skipped == 0.0) and np.all(g["ec_skip"] == 0.0)


# ------------------------------------------------------------------------ oracle, synthetic
@pytest.mark.parametrize("kwargs", [
    dict(),                                            # resample, defaults
    dict(ethresh=1e-6),                                # many skipped lines
    dict(cutoff=0.0, extent=20.0),                     # no fixed cutoff
    dict(resolution=8000.0),                           # constant-R (2-point interpolation)
    dict(nlines=200000, wnosamp=120),                  # dense: heavy co-adding
    dict(nlines=400000, wnosamp=60),                   # co-add groups longer than a warp
    dict(wnstep=0.25, wnosamp=360, wnlow=9000.0, wnhigh=9060.0),
    dict(nlines=300, wnlow=9000.0, wnhigh=9600.0),     # sparse: chunks span many outputs
    dict(nlines=4000, wnstep=0.1, wnosamp=240, wnlow=9000.0, wnhigh=9030.0),  # > 8 passes/line
])
@pytest.mark.parametrize("acc_mode", ["auto", "owner", "strided"])
def test_extinction_matches_oracle_synthetic(kwargs, acc_mode, monkeypatch):
    # "auto": chunk-owned kernel gathering from the output-stride Voigt table where the grid
    # allows it; "owner": the output-owned kernel on the same table; "strided": the generic
    # gather from the reference-layout table (csrc/lbl_kernels.cu).
    if acc_mode != "auto" and kwargs.get("resolution"):
        pytest.skip("constant-R grids have a single accumulate mode")
    if acc_mode == "strided":
        monkeypatch.setenv("PB200_ACC_MODE", "strided")
    if acc_mode == "owner":
        monkeypatch.setenv("PB200_ACC_KERNEL", "owner")
    case = helpers.synthetic_case(**kwargs)
    eng = _engine_for(case, profile="host")
    atm = case.atm
    temps = atm.temp
    dens = atm.d
    isoz = helpers.partition(case, temps).T
    resolution = case.spec.interpolate
    for add in (0, 1):
        want, wcnt = _oracle_units(case, temps, dens, isoz, case.iso_mol_index, 1, add,
                                   resolution)
        got, cnt = eng.extinction_batch(temps, dens, isoz, case.iso_mol_index, 1,
                                        case.ethresh, add, resolution, counters=True)
        assert got.shape == want.shape
        assert np.array_equal(cnt[:, :4], wcnt), "nadd/nskip/neval/sample counters differ"
        assert _peak_err(got, want) < TOL_PEAK
    eng.close()


def test_extinction_device_voigt_table_end_to_end():
    """Same as above but with the Voigt table built on the device (both kernels chained)."""
    case = helpers.synthetic_case(nlines=5000)
    eng = _engine_for(case, profile="device")
    temps, dens = case.atm.temp, case.atm.d
    isoz = helpers.partition(case, temps).T
    want, _ = _oracle_units(case, temps, dens, isoz, case.iso_mol_index, 1, 1, 0)
    got = eng.extinction_batch(temps, dens, isoz, case.iso_mol_index, 1, case.ethresh, 1, 0)
    assert _peak_err(got, want) < TOL_PEAK
    eng.close()


def test_multi_row_and_skipped_isotopes():
    """add=0 with two output rows (isotopes mapped to different species rows) and one
    isotope disabled through iso_iext=-1 (skip_mol, pyrat/extinction.py:165-168)."""
    case = helpers.synthetic_case(nlines=8000)
    eng = _engine_for(case, profile="host")
    temps, dens = case.atm.temp[:4], case.atm.d[:4]
    isoz = helpers.partition(case, temps).T
    iext = np.array([0, 1, -1, 1])
    want, wcnt = _oracle_units(case, temps, dens, isoz, iext, 2, 0, 0)
    got, cnt = eng.extinction_batch(temps, dens, isoz, iext, 2, case.ethresh, 0, 0,
                                    counters=True)
    assert got.shape == (4, 2, case.spec.nwave)
    assert np.array_equal(cnt[:, :4], wcnt)
    assert _peak_err(got, want) < TOL_PEAK
    eng.close()


def test_partition_tables_on_engine_and_shared_temperature_passes():
    """unit_isoz=NULL uses the engine's Z(T) tables; units sharing (T, Z) share one
    strengths pass (table mode: many pressures per temperature)."""
    case = helpers.synthetic_case(nlines=5000)
    eng = _engine_for(case, profile="host")
    eng.set_partition(case.db.temp, case.db.iso_pf)
    temps = np.repeat([500.0, 1500.5, 2500.25], 3)
    press = np.tile([1e-4, 1e-1, 10.0], 3)
    dens = case.atm.vmr[0] * press[:, None] * pc.bar / (pc.k * temps[:, None])
    isoz = helpers.partition(case, temps).T
    want, _ = _oracle_units(case, temps, dens, isoz, case.iso_mol_index, 1, 0, 0)
    got = eng.extinction_batch(temps, dens, None, case.iso_mol_index, 1, case.ethresh, 0, 0)
    assert _peak_err(got, want) < TOL_PEAK
    eng.close()


def test_edge_cases_empty_and_out_of_window_lines():
    case = helpers.synthetic_case(nlines=2000)
    from pyratbay_b200.engine import Engine
    eng = Engine(0)
    eng.set_grid(case.spec.wn, case.spec.own, case.spec.odivisors)
    eng.set_species(case.atm.mol_radius, case.atm.mol_mass, case.iso_atm_index,
                    case.iso_mass, case.iso_ratio)
    eng.set_voigt(case.lorentz, case.doppler, case.size, case.index, case.profile,
                  case.cutoff)
    temps, dens = case.atm.temp[:2], case.atm.d[:2]
    isoz = helpers.partition(case, temps).T
    # no lines at all
    eng.set_lines(np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0, int))
    out = eng.extinction_batch(temps, dens, isoz, case.iso_mol_index, 1, 1e-30, 1, 0)
    assert out.shape == (2, 1, case.spec.nwave) and np.all(out == 0.0)
    # every line outside the window
    eng.set_lines(np.array([10.0, 20.0]), np.ones(2), np.ones(2), np.zeros(2, int))
    assert eng.line_stats() == {"in_window": 0, "groups": 0, "nadd": 0}
    out = eng.extinction_batch(temps, dens, isoz, case.iso_mol_index, 1, 1e-30, 1, 0)
    assert np.all(out == 0.0)
    # lines on the window edges and duplicates on one fine sample
    own = case.spec.own
    lw = np.array([own[0], own[5], own[5], own[5] + 0.4 * case.spec.ownstep, own[-1]])
    eng.set_lines(lw, np.full(5, 100.0), np.full(5, 1e-6), np.zeros(5, int))
    c2 = helpers.Case(**{**case.__dict__, "lwn": lw, "elow": np.full(5, 100.0),
                         "gf": np.full(5, 1e-6), "isoid": np.zeros(5, int)})
    want, wcnt = _oracle_units(c2, temps, dens, isoz, case.iso_mol_index, 1, 1, 0)
    got, cnt = eng.extinction_batch(temps, dens, isoz, case.iso_mol_index, 1, 1e-30, 1, 0,
                                    counters=True)
    assert np.array_equal(cnt[:, :4], wcnt)
    assert _peak_err(got, want) < TOL_PEAK
    # zero units
    out = eng.extinction_batch(np.zeros(0), np.zeros((0, case.atm.nmol)),
                               np.zeros((0, 4)), case.iso_mol_index, 1, 1e-30, 1, 0)
    assert out.shape == (0, 1, case.spec.nwave)
    eng.close()


def test_error_behaviour():
    from pyratbay_b200.engine import Engine
    from pyratbay_b200._lib import PB200Error
    case = helpers.synthetic_case(nlines=2000)
    eng = Engine(0)
    with pytest.raises(PB200Error):   # lines before grid
        eng.set_lines(case.lwn, case.elow, case.gf, case.isoid)
    eng.set_grid(case.spec.wn, case.spec.own, case.spec.odivisors)
    eng.set_species(case.atm.mol_radius, case.atm.mol_mass, case.iso_atm_index,
                    case.iso_mass, case.iso_ratio)
    with pytest.raises(PB200Error):   # unsorted within an isotope
        eng.set_lines(case.lwn[::-1].copy(), case.elow, case.gf, case.isoid)
    with pytest.raises(PB200Error):   # batch before voigt/lines
        eng.extinction_batch(case.atm.temp, case.atm.d, None, case.iso_mol_index, 1, 1e-30,
                             1, 0)
    eng.close()


# ----------------------------------------------------------------------- table interpolation
def test_interp_ec_bit_exact_vs_oracle():
    from pyratbay_b200.engine import interp_ec, interp_ec_per_mol
    orc = helpers.oracle_module()
    for nwave in (1000, 777):       # even: 16-byte vector path, odd: scalar path
        _check_interp_bit_exact(interp_ec, interp_ec_per_mol, orc, nwave)


def _check_interp_bit_exact(interp_ec, interp_ec_per_mol, orc, nwave):
    rng = np.random.default_rng(3)
    nspec, ntemp, nlayers = 3, 7, 11
    table = rng.uniform(1e-30, 1e-18, (nspec, ntemp, nlayers, nwave))
    tgrid = np.linspace(300.0, 3000.0, ntemp)
    temp = rng.uniform(300.0, 3000.0, nlayers)
    temp[0], temp[1], temp[2] = 300.0, 3000.0, tgrid[3]   # grid nodes and ends
    dens = rng.uniform(1e8, 1e18, (nlayers, nspec))
    for fn_gpu, fn_cpu, shape in ((interp_ec, orc.interp_ec, (nlayers, nwave)),
                                  (interp_ec_per_mol, orc.interp_ec_per_mol,
                                   (nspec, nlayers, nwave))):
        for lay1, lay2 in ((0, nlayers), (3, 4), (2, 50)):
            a = rng.uniform(0.0, 1e-3, shape)
            b = a.copy()
            fn_gpu(a, table, tgrid, temp, dens, lay1, lay2)
            fn_cpu(b, table, tgrid, temp, dens, lay1, lay2)
            assert np.array_equal(a, b)


def test_line_sample_matches_reference():
    import pyratbay_b200 as pb
    g = helpers.golden("mock_line_sample.npz")
    ls = pb.Line_Sample(os.path.join(helpers.GOLDEN, "mock_opacity_file.npz"))
    temp, dens = g["temperature"], g["density"]
    np.testing.assert_allclose(ls.calc_cross_section(temp), g["cs"], rtol=1e-14)
    np.testing.assert_allclose(ls.calc_cross_section(temp, per_mol=True), g["cs_per_mol"],
                               rtol=1e-14)
    np.testing.assert_allclose(ls.calc_extinction_coefficient(temp, dens), g["ec"],
                               rtol=1e-14)
    np.testing.assert_allclose(ls.calc_extinction_coefficient(temp, dens, layer=20),
                               g["ec_layer"], rtol=1e-14)
    with pytest.raises(ValueError):
        ls.calc_cross_section(np.full(ls.nlayers, 5000.0))


def test_table_regridding_matches_reference():
    """pb200_regrid_table_dev (tools/tools.py:1026-1107 on the device) against the outputs of the
    unmodified reference's interpolate_opacity (tests/golden/make_golden_regrid.py): both axes
    with out-of-range and on-node samples, one axis only, wavenumber mask + thinning, the -230
    floor of zero / underflowing entries."""
    from pyratbay_b200 import io
    from pyratbay_b200.line_sampling import interpolate_opacity
    g = helpers.golden("mock_regrid.npz")
    path = os.path.join(helpers.GOLDEN, "mock_opacity_file.npz")
    _, temp, press, _wn = io.read_opacity(path, extract="arrays")
    tol = dict(rtol=1e-12, atol=0)
    np.testing.assert_allclose(interpolate_opacity(path, g["t_both"], g["p_both"]), g["cs_both"], **tol)
    np.testing.assert_allclose(interpolate_opacity(path, g["t_only"], press, g["mask"], 3),
                               g["cs_t_only"], **tol)
    np.testing.assert_allclose(interpolate_opacity(path, temp, g["p_only"]), g["cs_p_only"], **tol)
    same = interpolate_opacity(path, temp, press)
    assert np.array_equal(same, io.read_opacity(path, extract="opacity"))


def test_table_regridding_floor_of_zero_entries(tmp_path):
    from pyratbay_b200 import io
    from pyratbay_b200.line_sampling import interpolate_opacity
    g = helpers.golden("mock_regrid.npz")
    path = os.path.join(helpers.GOLDEN, "mock_opacity_file.npz")
    _, species, temp, press, wn, _ = io.read_opacity(path, extract="all")
    zpath = str(tmp_path / "zero_table.npz")
    io.write_opacity(zpath, species, temp, press, wn, g["zero_table"])
    got = interpolate_opacity(zpath, g["t_both"], g["p_both"])
    want = g["cs_zero"]
    assert np.all(np.isfinite(got)) and got.min() >= 0
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-300)


def test_line_sample_regridded_axes_and_str():
    """Line_Sample built on other pressure / temperature axes than the file's (device
    re-gridding), its extinction and its text form against the reference."""
    import pyratbay_b200 as pb
    g = helpers.golden("mock_regrid.npz")
    path = os.path.join(helpers.GOLDEN, "mock_opacity_file.npz")
    ls = pb.Line_Sample(path, pressure=g["p_both"][1:6], temperature=g["t_both"][1:6])
    np.testing.assert_allclose(ls.cs_table, g["ls_cs_table"], rtol=1e-12)
    dens = np.full((5, 1), 1.0e12)
    np.testing.assert_allclose(ls.calc_extinction_coefficient(g["ls_temp"], dens), g["ls_ec"],
                               rtol=1e-12)
    assert str(ls).replace(path, "FILE") == str(g["ls_str"])


def test_line_sample_from_device_table_without_file_round_trip(tmp_path):
    """compute_opacity leaves the table in HBM; Line_Sample(tables=[pyrat.ex]) consumes it there
    and gives the bits of a Line_Sample built from the .npz file; device_out=True queues the
    interpolation without synchronising and is bit-exact with the oracle's interp_ec."""
    import torch
    import pyratbay_b200 as pb
    from pyratbay_b200 import tli as ptli, workloads
    from pyratbay_b200.pyrat import Pyrat
    w = workloads.table_workload(100_000, ntemp=5, nlayers=6, nwave=1500, wl_low_um=1.0,
                                 wl_high_um=1.2)
    tli_path = str(tmp_path / "lines.tli")
    wn, elow, gf, iso, counts = w.make_lines()
    ptli.write_tli(tli_path, [w.db], [{"wn": wn, "elow": elow, "gf": gf, "iso_id": iso,
                                       "n_lines_iso": counts}], w.inputs["wnlow"], w.inputs["wnhigh"])
    cs = str(tmp_path / "table.npz")
    pyrat = Pyrat(dict(w.inputs, tlifile=[tli_path], sampled_cs=[cs]), atm=w.atm, device=0)
    pyrat.compute_opacity()
    ex = pyrat.ex
    from_dev = pb.Line_Sample(tables=[ex])
    from_file = pb.Line_Sample(cs)
    assert from_dev.cs_table_device.data_ptr() != ex.etable_dev.data_ptr()   # its own copy
    assert np.array_equal(from_dev.cs_table, from_file.cs_table)
    assert np.array_equal(from_dev.cs_table[0], ex.etable)
    temp = np.linspace(350.0, 850.0, 6)
    dens = np.geomspace(1e10, 1e18, 6)[:, None]
    a = from_dev.calc_extinction_coefficient(temp, dens)
    b = from_file.calc_extinction_coefficient(temp, dens)
    assert np.array_equal(a, b)
    orc = helpers.oracle_module()
    want = np.zeros((6, 1500))
    orc.interp_ec(want, from_file.cs_table, from_file.temp, temp, dens, 0, 6)
    assert np.array_equal(a, want)
    # asynchronous, device-resident result: same bits, many calls back to back
    for _ in range(20):
        d = from_dev.calc_extinction_coefficient(temp, dens, device_out=True)
    assert isinstance(d, torch.Tensor) and d.is_cuda
    assert np.array_equal(d.cpu().numpy(), want)
    per = from_dev.calc_extinction_coefficient(temp, dens, per_mol=True, device_out=True)
    wantm = np.zeros((1, 6, 1500))
    orc.interp_ec_per_mol(wantm, from_file.cs_table, from_file.temp, temp, dens, 0, 6)
    assert np.array_equal(per.cpu().numpy(), wantm)
    # a layer range leaves the other rows at zero, as the reference's zeroed array
    part = from_dev.calc_extinction_coefficient(temp, dens, layer=(2, 4))
    wantp = np.zeros((6, 1500))
    orc.interp_ec(wantp, from_file.cs_table, from_file.temp, temp, dens, 2, 4)
    assert np.array_equal(part, wantp)


# ------------------------------------------------------------- full-size configs[1] checks
@pytest.fixture(scope="module")
def full_size_engine():
    """BASELINE.json configs[1]: 1e6 synthetic H2O lines, 81 layers, 0.5-5 um, add=1."""
    from pyratbay_b200 import workloads
    from pyratbay_b200.engine import Engine
    from pyratbay_b200.voigt import Voigt
    w = workloads.forward_model_workload(1_000_000, 81)
    eng = Engine(0)
    eng.set_grid(w.spec.wn, w.spec.own, w.spec.odivisors)
    eng.set_species(w.atm.mol_radius, w.atm.mol_mass, w.iso_atm_index, w.db.iso_mass,
                    w.db.iso_ratio)
    eng.set_lines(w.wn, w.elow, w.gf, w.isoid)
    voigt = Voigt(w.spec, w.atm, w.iso_atm_index, eng)
    temps = w.atm.temp
    isoz = workloads.partition(w.db, temps)
    yield w, eng, voigt, temps, isoz
    eng.close()


def test_full_size_spot_check_against_oracle(full_size_engine):
    """Three of the 81 layers of the benchmark workload, every line, against the CPU oracle
    (low, middle and high pressure: narrow, mixed and Lorentz-dominated profiles)."""
    w, eng, voigt, temps, isoz = full_size_engine
    orc = helpers.oracle_module()
    got, cnt = eng.extinction_batch(temps, w.atm.d, isoz, w.iso_mol_index, 1, 1e-30, 1, 0,
                                    counters=True)
    assert np.all(np.isfinite(got)) and np.all(got >= 0)
    profile = voigt.profile     # device table copied back; the oracle reads the same table
    for layer in (3, 40, 80):
        ext = np.zeros((1, w.spec.nwave))
        ocnt = np.zeros(4, np.int64)
        orc.extinction(ext, profile, voigt.size, voigt.index, voigt.lorentz, voigt.doppler,
                       w.spec.wn, w.spec.own, w.spec.odivisors, w.atm.d[layer],
                       w.atm.mol_radius, w.atm.mol_mass, w.iso_atm_index, w.db.iso_mass,
                       w.db.iso_ratio, isoz[layer], w.iso_mol_index, w.wn, w.elow, w.gf,
                       w.isoid, voigt.cutoff, 1e-30, temps[layer], 0, 1, 0, counters=ocnt)
        assert np.array_equal(cnt[layer, :4], ocnt)
        assert _peak_err(got[layer], ext) < TOL_PEAK


def test_full_size_exact_linearity_and_isotope_additivity(full_size_engine):
    """Size-independent properties at the benchmark size.
    (1) Doubling every gf doubles every strength exactly (power of two), leaves the ethresh
        selection unchanged and so must double the output BIT FOR BIT.
    (2) Co-add groups never span isotopes, so the extinction of all isotopes equals the sum
        of the four single-isotope runs (summation order aside)."""
    w, eng, voigt, temps, isoz = full_size_engine
    args = (temps, w.atm.d, isoz)
    base = eng.extinction_batch(*args, w.iso_mol_index, 1, 1e-30, 1, 0)
    parts = np.zeros_like(base)
    for i in range(w.db.niso):
        iext = np.full(w.db.niso, -1)
        iext[i] = 0
        parts += eng.extinction_batch(*args, iext, 1, 1e-30, 1, 0)
    assert _peak_err(parts, base) < 1e-12
    eng.set_lines(w.wn, w.elow, 2.0 * w.gf, w.isoid)
    doubled = eng.extinction_batch(*args, w.iso_mol_index, 1, 1e-30, 1, 0)
    eng.set_lines(w.wn, w.elow, w.gf, w.isoid)
    assert np.array_equal(doubled, 2.0 * base)
    # checksum of checksums: per-layer sums add up to the grand total
    assert abs(np.sum(np.sum(base, axis=-1)) / np.sum(base) - 1.0) < 1e-13


# --------------------------------------------------------------- optical depth (next tier)
@pytest.mark.parametrize("name", ["emission", "emission_max", "transit", "transit_max"])
def test_optical_depth_matches_oracle_and_reference(name):
    from pyratbay_b200.optic_depth import optical_depth, transit_path, _path_matrix
    orc = helpers.oracle_module()
    g = helpers.golden("mock_optical_depth.npz")
    ec = np.ascontiguousarray(g["ec"] * 3e4)
    nlayers, nwave = ec.shape
    kw = {"emission": dict(rt_path="emission"),
          "emission_max": dict(rt_path="emission", maxdepth=10.0, itop=3, ibottom=45),
          "transit": dict(rt_path="transit"),
          "transit_max": dict(rt_path="transit", maxdepth=10.0, itop=2, ibottom=48)}[name]
    raypath, depth, ideep, dclear, iclear = optical_depth(extinction=ec, radius=g["radius"],
                                                          **kw)
    assert dclear is None and iclear is None
    np.testing.assert_allclose(depth, g[name + "_depth"], rtol=1e-13, atol=0)
    assert np.array_equal(ideep, g[name + "_ideep"])
    # bit-exact against the strict-IEEE oracle
    itop = kw.get("itop", 0)
    ibottom = kw.get("ibottom", nlayers)
    maxdepth = kw.get("maxdepth", np.inf)
    want = np.zeros((nlayers, nwave))
    wdeep = np.zeros(nwave, np.int32)
    if kw["rt_path"] == "transit":
        orc.transit_optical_depth(want, wdeep, ec, _path_matrix(transit_path(g["radius"], itop),
                                                                 nlayers), maxdepth, itop, ibottom)
    else:
        orc.plane_parallel_optical_depth(want, wdeep, ec, -np.ediff1d(g["radius"]), maxdepth,
                                         itop, ibottom)
    assert np.array_equal(depth, want) and np.array_equal(ideep, wdeep)


def test_optical_depth_patchy_and_errors():
    from pyratbay_b200.optic_depth import optical_depth
    g = helpers.golden("mock_optical_depth.npz")
    ec = g["ec"] * 3e4
    rp, depth, ideep, dclear, iclear = optical_depth(
        "transit", ec, radius=g["radius"], maxdepth=10.0,
        extinction_cloudy=np.full_like(ec, 2e-9))
    np.testing.assert_allclose(depth, g["patchy_depth"], rtol=1e-13)
    np.testing.assert_allclose(dclear, g["patchy_depth_clear"], rtol=1e-13)
    assert np.array_equal(ideep, g["patchy_ideep"])
    assert np.array_equal(iclear, g["patchy_ideep_clear"])
    with pytest.raises(ValueError):
        optical_depth("sideways", ec, radius=g["radius"])
    with pytest.raises(ValueError):
        optical_depth("transit", ec)


def test_line_sample_isotope_ratios_match_reference(tmp_path):
    """Two tables labelled as isotopologues with one free and one filler ratio
    (line_sampling.py:142-229), extinction before and after updating the parameter."""
    import shutil
    import pyratbay_b200 as pb
    g = helpers.golden("mock_line_sample_iso.npz")
    g1 = helpers.golden("mock_line_sample.npz")
    src = os.path.join(helpers.GOLDEN, "mock_opacity_file.npz")
    a, b = str(tmp_path / "cs_H2O_161.npz"), str(tmp_path / "cs_H2O_181.npz")
    shutil.copy(src, a)
    shutil.copy(src, b)
    ls = pb.Line_Sample([a, b], isotope_ratios="161 main fill_heavy\n181 heavy -2.5")
    assert list(ls.species) == list(g["species"]) and list(ls.pnames) == list(g["pnames"])
    assert np.array_equal(ls.pars, g["pars"])
    temp = g1["temperature"]
    dens = np.tile(g1["density"], (1, 2))
    np.testing.assert_allclose(ls.calc_extinction_coefficient(temp, dens), g["ec"], rtol=1e-14)
    np.testing.assert_allclose(ls.calc_extinction_coefficient(temp, dens, pars=[-3.0]),
                               g["ec_pars"], rtol=1e-14)
    np.testing.assert_allclose(ls.iso_ratios, g["iso_ratios_after"], rtol=0, atol=0)
    np.testing.assert_allclose(ls.calc_cross_section(temp, per_mol=True), g["cs_per_mol"],
                               rtol=1e-14)


def test_two_species_through_line_by_line_api(tmp_path):
    """Two single-database TLI files (the reference offsets isotope ids per file,
    line_by_line.py:119-121): co-added extinction, per-species get_ec and skip_mol, through
    the Pyrat-shaped API, against the oracle fed with the same concatenated arrays."""
    import pyratbay_b200 as pb
    from pyratbay_b200 import tli as ptli
    orc = helpers.oracle_module()
    temp, z = ptli.h2o_partition_table()
    db_a = ptli.Database("Synthetic H2O", "H2O", temp, ["116", "118"], [18.01056, 20.01481],
                         [0.9973, 0.0020], z[:2])
    db_b = ptli.Database("Synthetic CO", "CO", temp, ["26", "36", "28"],
                         [27.9949, 28.9983, 29.9992], [0.9865, 0.0111, 0.0020], z[1:4] * 0.31)
    files = []
    for k, (db, n) in enumerate([(db_a, 6000), (db_b, 4000)]):
        fr = tuple(np.array([0.8, 0.2]) if db.niso == 2 else np.array([0.7, 0.2, 0.1]))
        wn, elow, gf, iso, counts = ptli.synthetic_lines(n, 4990.0, 5215.0, fractions=fr,
                                                         seed=10 + k)
        path = str(tmp_path / f"syn_{k}.tli")
        ptli.write_tli(path, [db], [{"wn": wn, "elow": elow, "gf": gf, "iso_id": iso,
                                     "n_lines_iso": counts}], 4990.0, 5215.0)
        files.append(path)
    a = helpers.golden("mock_atmosphere.npz")
    nlayers = 7
    from pyratbay_b200 import atmosphere as pa
    atm = pa.Atmosphere(pa.pressure(1e-5, 50.0, nlayers), np.linspace(600.0, 2400.0, nlayers),
                        np.tile(a["vmr"][0], (nlayers, 1)), [str(s) for s in a["species"]])
    inputs = dict(tlifile=files, wnlow=5000.0, wnhigh=5200.0, wnstep=1.0, wnosamp=720,
                  voigt_extent=50.0, voigt_nlor=30, voigt_ndop=12, tmin=300.0, tmax=3000.0,
                  verb=0)
    pyrat = pb.Pyrat(inputs, atm=atm)
    lbl, spec, voigt = pyrat.lbl, pyrat.spec, pyrat.voigt
    assert list(lbl.species) == ["CO", "H2O"] and lbl.nspec == 2 and lbl.niso == 5
    assert list(lbl.iso_mol_index) == [1, 1, 0, 0, 0]
    profile = voigt.profile
    isoz = lbl.partition(atm.temp).T

    def oracle(layer, add, iext):
        ext = np.zeros((lbl.nspec, spec.nwave))
        orc.extinction(ext, profile, voigt.size, voigt.index, voigt.lorentz, voigt.doppler,
                       spec.wn, spec.own, spec.odivisors, atm.d[layer], atm.mol_radius,
                       atm.mol_mass, lbl.iso_atm_index, lbl.iso_mass, lbl.iso_ratio,
                       isoz[layer], iext, lbl.wn, lbl.elow, lbl.gf, lbl.isoid, voigt.cutoff,
                       lbl.ethresh, atm.temp[layer], 0, int(add), 0)
        return ext

    ec = pyrat.calc_lbl_extinction()
    for layer in range(nlayers):
        assert _peak_err(ec[layer], oracle(layer, True, lbl.iso_mol_index)[0]) < TOL_PEAK
    per_species, label = pyrat.get_ec(3)
    want = oracle(3, False, lbl.iso_mol_index) * atm.d[3, lbl.mol_index][:, None]
    assert label == ["CO", "H2O"] and _peak_err(per_species, want) < TOL_PEAK
    skipped = np.copy(pyrat.calc_lbl_extinction(skip_mol=["CO"]))
    iext = np.where(lbl.iso_mol_index == 0, -1, lbl.iso_mol_index)
    for layer in (0, 6):
        assert _peak_err(skipped[layer], oracle(layer, True, iext)[0]) < TOL_PEAK
    with pytest.raises(ValueError):   # tables are single-species (pyrat/extinction.py:57-62)
        pyrat.ex.tmin, pyrat.ex.tmax, pyrat.ex.tstep = 300.0, 3000.0, 300.0
        pyrat.ex.sampled_cs = [str(tmp_path / "t.npz")]
        pyrat.compute_opacity()


@pytest.mark.parametrize("kwargs", [
    dict(nlines=20000),
    dict(nlines=200000, wnosamp=120),                       # dense: long co-add chains
    dict(nlines=60000, wnstep=0.25, wnosamp=360, wnlow=9000.0, wnhigh=9060.0),
])
def test_device_line_preprocessing_matches_host(kwargs, monkeypatch):
    """pb200_engine_set_lines on the device (segment-parallel co-add walk, csrc/preprocess.cu)
    against the sequential host walk: identical groups, hence bit-identical extinction."""
    case = helpers.synthetic_case(**kwargs)
    temps, dens = case.atm.temp, case.atm.d
    isoz = helpers.partition(case, temps).T
    results = {}
    for mode in ("host", "device"):
        monkeypatch.setenv("PB200_SETLINES", mode)
        eng = _engine_for(case, profile="host")
        results[mode] = (eng.line_stats(),
                         eng.extinction_batch(temps, dens, isoz, case.iso_mol_index, 1,
                                              case.ethresh, 1, 0, counters=True))
        eng.close()
    assert results["host"][0] == results["device"][0]
    assert results["host"][0]["nadd"] > 0
    assert np.array_equal(results["host"][1][0], results["device"][1][0])
    assert np.array_equal(results["host"][1][1], results["device"][1][1])


def test_device_line_preprocessing_edge_cases(monkeypatch):
    """Window edges, duplicates, blocks entirely outside the window, a single line."""
    case = helpers.synthetic_case(nlines=2000)
    own, step = case.spec.own, case.spec.ownstep
    temps, dens = case.atm.temp[:2], case.atm.d[:2]
    isoz = helpers.partition(case, temps).T
    blocks = [
        # isotope 0: below the window, on its edge, duplicates and near-duplicates
        np.array([own[0] - 3.0, own[0] - 1e-9, own[0], own[5], own[5], own[5] + 0.4 * step,
                  own[5] + 0.9 * step, own[5] + 1.1 * step, own[9] + 0.5 * step, own[-1],
                  own[-1] + 1e-9, own[-1] + 2.0]),
        np.array([own[0] - 9.0, own[0] - 8.0]),                    # isotope 1: all below
        np.array([own[100] + 0.49 * step]),                        # isotope 2: one line
        np.array([own[-1] + 1.0, own[-1] + 2.0, own[-1] + 2.0]),   # isotope 3: all above
    ]
    lw = np.concatenate(blocks)
    iso = np.concatenate([np.full(len(b), i) for i, b in enumerate(blocks)])
    elow, gf = np.full(len(lw), 100.0), np.full(len(lw), 1e-6)
    out = {}
    for mode in ("host", "device"):
        monkeypatch.setenv("PB200_SETLINES", mode)
        eng = _engine_for(helpers.Case(**{**case.__dict__, "lwn": lw, "elow": elow, "gf": gf,
                                          "isoid": iso}), profile="host")
        out[mode] = (eng.line_stats(), eng.extinction_batch(
            temps, dens, isoz, case.iso_mol_index, 1, 1e-30, 1, 0, counters=True))
        eng.close()
    assert out["host"][0] == out["device"][0]
    assert np.array_equal(out["host"][1][0], out["device"][1][0])
    assert np.array_equal(out["host"][1][1], out["device"][1][1])


def test_exact_quotient_and_threshold_index_selftest():
    """The two exactness shortcuts of the accumulate kernel's prepare step (csrc/common.cuh):
    idwn = (int)((wn - own0)/dwnstep) (_extcoeff.c:275) from a host-rounded reciprocal with
    FMA corrections must equal the IEEE division bit for bit, and the Doppler threshold table
    must give the index of the reference's nearest-sample search (:278) for every width --
    including widths exactly on and one ulp below a threshold."""
    from pyratbay_b200.engine import selftest_exact
    from pyratbay_b200.spectrum import Spectrum
    spec = Spectrum(wnlow=2000.0, wnhigh=20000.0, wnstep=1.0, wnosamp=2160)
    steps = spec.ownstep * np.asarray(spec.odivisors, float)
    steps = np.concatenate([steps, [1.0 / 2520, 0.25 / 360, 1.0 / 3, 1e-3 * (1 + 2.0**-52)]])
    doppler = np.logspace(np.log10(2.3e-3), np.log10(0.27), 50)
    for seed in (1, 2):
        bad_q, bad_i = selftest_exact(steps, doppler, n=1 << 25, seed=seed)
        assert bad_q == 0 and bad_i == 0
