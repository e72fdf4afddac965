"""CPU tests of the host-side logic and of the C-ABI library's surface (no GPU needed)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import helpers
from pyratbay_b200 import constants as pc


# ------------------------------------------------------------------------------- TLI files
def test_read_tli_file_matches_reference_windows():
    from pyratbay_b200.tli import read_tli_file
    g = helpers.golden("tli_window_cases.npz")
    tli = os.path.join(helpers.GOLDEN, "mock_hitran_h2o.tli")
    k = 0
    while f"w{k}_range" in g:
        lo, hi = g[f"w{k}_range"]
        dbs, wn, gf, elow, iso = read_tli_file(tli, lo, hi)
        assert np.array_equal(wn, g[f"w{k}_wn"]), k
        assert np.array_equal(gf, g[f"w{k}_gf"])
        assert np.array_equal(elow, g[f"w{k}_elow"])
        assert np.array_equal(iso, g[f"w{k}_iso"])
        k += 1
    assert k == 7
    db = dbs[0]
    # header facts printed by the reference's tests/test_tli.py:16-38
    assert db.name == "HITRAN H2O" and db.molname == "H2O"
    assert db.niso == 4 and db.ntemp == 1201
    assert [str(x) for x in db.iso_name] == ["116", "118", "117", "126"]
    np.testing.assert_allclose(db.iso_mass, [18.0106, 20.0148, 19.0148, 19.0167], atol=1e-4)
    dbs, wn, gf, elow, iso = read_tli_file(tli, 0.0, 1e5)
    assert len(wn) == 888
    assert list(np.bincount(iso)) == [672, 148, 62, 6]
    for i in range(4):
        assert np.all(np.diff(wn[iso == i]) >= 0)


def test_tli_write_read_round_trip_and_errors(tmp_path):
    from pyratbay_b200 import tli as ptli
    path = str(tmp_path / "syn.tli")
    db = ptli.make_synthetic_tli(path, 5000, 4000.0, 4100.0, seed=3)
    dbs, wn, gf, elow, iso = ptli.read_tli_file(path, 0.0, 1e6)
    w2, e2, g2, i2, counts = ptli.synthetic_lines(5000, 4000.0, 4100.0, seed=3)
    assert np.array_equal(wn, w2) and np.array_equal(elow, e2) and np.array_equal(gf, g2)
    assert np.array_equal(iso, i2) and list(counts) == [3750, 750, 350, 150]
    assert dbs[0].niso == 4 and np.array_equal(dbs[0].iso_pf, db.iso_pf)
    # window extraction keeps per-isotope blocks
    _, wn, _, _, iso = ptli.read_tli_file(path, 4040.0, 4050.0)
    assert np.all((wn >= 4040.0) & (wn <= 4050.0)) and np.all(np.diff(iso) >= 0)
    # empty window
    _, wn, _, _, _ = ptli.read_tli_file(path, 10.0, 20.0)
    assert len(wn) == 0
    # truncated file
    bad = str(tmp_path / "bad.tli")
    open(bad, "wb").write(open(path, "rb").read()[:-7])
    with pytest.raises(ValueError):
        ptli.read_tli_file(bad, 0.0, 1e6)


# ----------------------------------------------------------------------------- grids, sizing
def test_spectrum_grids_match_reference():
    from pyratbay_b200.spectrum import Spectrum, constant_resolution_spectrum
    g = helpers.golden("mock_opacity_table.npz")
    spec = Spectrum(wl_low=1.00 * pc.um, wl_high=1.01 * pc.um, wnstep=1.0, wnosamp=2160)
    assert np.array_equal(spec.wn, g["wn"])
    assert spec.own[0] == float(g["own0"]) and spec.ownstep == float(g["ownstep"])
    assert spec.onwave == int(g["onwave"])
    assert np.array_equal(spec.odivisors, g["odivisors"])
    # default oversampling: first highly composite h with wnstep/h <= 4e-4 (spectrum.py:192-195)
    assert Spectrum(wnlow=100.0, wnhigh=200.0, wnstep=1.0).wnosamp == 2520
    gR = helpers.golden("mock_opacity_table_R.npz")
    specR = Spectrum(wl_low=1.00 * pc.um, wl_high=1.01 * pc.um, wnstep=1.0, wnosamp=2160,
                     resolution=15000.0)
    assert np.array_equal(specR.wn, gR["wn"]) and specR.interpolate
    # docstring example of spec_tools.py:483-492
    wl = constant_resolution_spectrum(0.5, 4.0, 5.5)
    np.testing.assert_allclose(wl[:4], [0.5, 0.6, 0.72, 0.864])
    assert len(wl) == 12
    with pytest.raises(ValueError):
        Spectrum(wnlow=200.0, wnhigh=100.0, wnstep=1.0)
    with pytest.raises(ValueError):
        Spectrum(wnlow=100.0, wnhigh=200.0)


def test_broadening_known_answers():
    from pyratbay_b200 import broadening as b
    # docstring examples of broadening.py:394-407 and :458-471 (tests/test_broadening.py:130-155)
    dmin, lmin = b.min_widths(100.0, 3000.0, 1.0 / (10.0 * pc.um), 18.015, 1.6 * pc.A, 1e-5)
    assert f"{dmin:.2e}" == "8.44e-04" and f"{lmin:.2e}" == "2.21e-07"
    dmax, lmax = b.max_widths(100.0, 3000.0, 1.0 / (1.0 * pc.um), 18.015, 1.6 * pc.A, 100.0)
    assert f"{dmax:.2e}" == "4.62e-02" and f"{lmax:.2e}" == "1.21e+01"


def test_voigt_sizing_matches_reference():
    g = helpers.golden("mock_voigt.npz")
    case = helpers.mock_case(with_profile=False)
    v = case.voigt
    assert np.array_equal(v.lorentz, g["lorentz"]) and np.array_equal(v.doppler, g["doppler"])
    assert v.profile_len == int(g["profile_len"])
    # skipped profiles are 0 before the grid call and never in column 0
    assert np.all(v.size[:, 0] > 0)
    computed = v.size > 0
    assert np.array_equal(v.size[computed], g["size"][computed])
    from pyratbay_b200._lib import PB200Error
    with pytest.raises(PB200Error):
        v.profile     # no engine -> no table, never a CPU fallback


def test_atmosphere_and_config(tmp_path):
    from pyratbay_b200 import atmosphere as pa
    from pyratbay_b200 import tools as pt
    atm = helpers.mock_atmosphere()
    a = helpers.golden("mock_atmosphere.npz")
    assert np.array_equal(atm.mol_mass, a["mol_mass"])
    assert np.array_equal(atm.mol_radius, a["mol_radius"])
    assert np.array_equal(atm.d, a["d"])
    # .atm round trip
    atmfile = tmp_path / "uniform.atm"
    with open(atmfile, "w") as f:
        f.write("# test\n@PRESSURE\nbar\n@TEMPERATURE\nkelvin\n@ABUNDANCE\nvolume\n"
                "@SPECIES\nH2  He  H2O\n\n@DATA\n")
        for p, t in zip(atm.press[:5], atm.temp[:5]):
            f.write(f"{p:.6e} {t:.3f} 8.5e-01 1.49e-01 4.0e-04\n")
    species, press, temp, vmr = pa.read_atm(str(atmfile))
    assert species == ["H2", "He", "H2O"] and vmr.shape == (5, 3)
    np.testing.assert_allclose(press, atm.press[:5], rtol=1e-6)
    cfg = tmp_path / "opacity.cfg"
    cfg.write_text("[pyrat]\nrunmode = opacity\nlogfile = out/table.log\n"
                   f"atmfile = {atmfile}\ntlifile = a.tli\nwl_low = 1.1 um\nwl_high = 1.7 um\n"
                   "wnstep = 1.0\nwnosamp = 2160\nvoigt_extent = 100.0\ntmin = 300\n"
                   "tmax = 3000\ntstep = 300\nncpu = 7\nverb = 2\n")
    args = pt.parse(str(cfg))
    assert args.runmode == "opacity" and args.wnosamp == 2160 and args.voigt_extent == 100.0
    assert abs(args.wl_low - 1.1e-4) < 1e-18 and args.ethresh == 1e-30
    assert args.voigt_cutoff == 25.0 and args.voigt_ndop == 50 and args.voigt_nlor == 100
    assert args.sampled_cs[0].endswith("out/table.npz")      # parser.py:695-696
    assert os.path.isabs(args.tlifile[0])
    with pytest.raises(ValueError):
        pt.parse(str(tmp_path / "missing.cfg"))


def test_opacity_file_round_trip(tmp_path):
    from pyratbay_b200 import io
    rng = np.random.default_rng(0)
    temp, press = np.linspace(300, 3000, 4), np.logspace(-6, 2, 5)
    wn = np.linspace(1000, 1010, 11)
    table = rng.uniform(size=(4, 5, 11))
    path = str(tmp_path / "t.npz")
    io.write_opacity(path, "H2O", temp, press, wn, table)
    with np.load(path, allow_pickle=True) as f:     # layout of io.py:594-606
        assert sorted(f.files) == ["opacity", "pressure", "species", "temperature", "units",
                                   "wavenumber"]
        assert list(f["species"]) == ["H2O"]
        assert f["units"].item()["pressure"] == "bar"
    units, species, t2, p2, w2, tab2 = io.read_opacity(path)
    assert species == "H2O" and np.array_equal(tab2, table) and np.array_equal(p2, press)
    assert io.read_opacity(path, extract="opacity").shape == (4, 5, 11)
    assert io.read_opacity(path, extract="arrays")[0] == "H2O"
    with pytest.raises(ValueError):
        io.write_opacity(path, ["H2O"], temp, press, wn, table)
    # the reference's own file reads back identically
    _, sp, t, p, w, tab = io.read_opacity(os.path.join(helpers.GOLDEN, "mock_opacity_file.npz"))
    g = helpers.golden("mock_opacity_table.npz")
    assert sp == "H2O" and np.array_equal(tab, g["etable"]) and np.array_equal(w, g["wn"])


def test_regridding_brackets():
    """Host part of the table re-gridding (the kernel takes brackets and weights): piecewise
    linear with edge values outside the nodes, exact rows at the nodes (interp1d 'slinear' with
    fill_value=(first, last), tools/tools.py:1083-1105)."""
    from pyratbay_b200.line_sampling import _brackets, _needs_resampling
    nodes = np.array([300.0, 600.0, 900.0, 1500.0])
    lo, hi, f = _brackets(nodes, [100.0, 300.0, 450.0, 900.0, 1499.0, 1500.0, 2000.0])
    assert list(lo) == [0, 0, 0, 2, 2, 3, 3] and list(hi) == [0, 0, 1, 2, 3, 3, 3]
    np.testing.assert_allclose(f, [0, 0, 0.5, 0, 599.0 / 600.0, 0, 0], rtol=1e-15)
    assert not _needs_resampling(nodes, None) and not _needs_resampling(nodes, nodes * 1.005)
    assert _needs_resampling(nodes, nodes[:3]) and _needs_resampling(nodes, nodes * 1.02)


# --------------------------------------------------------------------------- C-ABI surface
def test_c_abi_library_exports_every_declared_symbol():
    from pyratbay_b200 import _lib, build
    build.build_library()
    header = open(os.path.join(helpers.ROOT, "include", "pb200_lbl.h")).read()
    declared = sorted(set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 22
    assert sorted(_lib.EXPORTS) == declared
    lib = ctypes.CDLL(_lib.lib_path())
    for name in declared:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path()],
                         capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", out), name
    assert b"sm_100a" in _lib.load().pb200_version()


def test_product_path_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pyratbay_b200 import _lib
    from pyratbay_b200.engine import Engine, voigt_grid, interp_ec
    assert _lib.device_count() == 0
    with pytest.raises(_lib.PB200Error):
        Engine(0)
    with pytest.raises(_lib.PB200Error):
        voigt_grid(np.zeros(10), np.ones((1, 1), np.int64), np.zeros((1, 1), np.int64),
                   [1e-3], [1e-2], 1e-3)
    with pytest.raises(_lib.PB200Error):
        interp_ec(np.zeros((2, 3)), np.zeros((1, 2, 2, 3)), [1.0, 2.0], [1.5, 1.5],
                  np.ones((2, 1)), 0, 2)
    # raw ABI: status code and message
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.pb200_engine_create(ctypes.c_int(0), ctypes.byref(h)) == -2  # PB200_ENODEVICE
    assert b"no CPU fallback" in lib.pb200_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(helpers.ROOT, "pyratbay_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "lbl_oracle" not in text, f
                assert "/root/reference" not in text, f


# ------------------------------------------------------------------------- multi-GPU logic
def test_partition_units():
    from pyratbay_b200.parallel import partition_units
    n = 1020
    for world in (1, 2, 3, 8):
        parts = [partition_units(n, r, world) for r in range(world)]
        assert sorted(np.concatenate(parts)) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    cost = np.tile(np.logspace(0, 1.5, 51), 20)
    parts = [partition_units(n, r, 8, cost) for r in range(8)]
    assert sorted(np.concatenate(parts)) == list(range(n))
    loads = [cost[p].sum() for p in parts]
    assert max(loads) / min(loads) < 1.02


def _gloo_worker(rank, world, port, tmp):
    import torch.distributed as dist
    sys.path.insert(0, helpers.ROOT)
    from pyratbay_b200 import parallel
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world)
    n_units, nwave = 23, 17
    full = np.arange(n_units * nwave, dtype=np.double).reshape(n_units, nwave) ** 1.5
    mine = parallel.partition_units(n_units, rank, world)
    table = np.zeros((n_units, nwave))
    table[mine] = full[mine]          # "compute" only this rank's units
    parallel.assemble_rows(table, mine)
    np.save(os.path.join(tmp, f"table_{rank}.npy"), table)
    dist.destroy_process_group()


def test_table_assembly_across_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as tmp_mp
    port = 29500 + os.getpid() % 2000
    tmp_mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    full = np.arange(23 * 17, dtype=np.double).reshape(23, 17) ** 1.5
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"table_{r}.npy"), full)


def _gloo_assembler_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, helpers.ROOT)
    from pyratbay_b200 import parallel
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world)
    n_units, nwave = 23, 17                       # 12 + 11 units: slots padded to 12 rows
    full = np.arange(n_units * nwave, dtype=np.double).reshape(n_units, nwave) ** 1.5
    cost = 1.0 + (np.arange(n_units) % 5)
    owners = parallel.unit_owners(n_units, world, cost, equal_counts=True)
    asm = parallel.TableAssembler(n_units, nwave, owners, rank, device=None, nchunks=3)
    seen = []
    for c, units, _ptr in asm.chunks():           # "compute" this rank's rows chunk by chunk
        lo = asm.bounds[c]
        asm.local[lo:lo + len(units)] = torch.from_numpy(full[units])
        seen.append(units)
        asm.publish(c)
    table = asm.finish().numpy()
    assert np.array_equal(np.concatenate(seen), owners[rank])
    np.save(os.path.join(tmp, f"asm_{rank}.npy"), table)
    dist.destroy_process_group()


def test_chunked_device_style_assembly_two_ranks_gloo(tmp_path):
    """TableAssembler (the product path of compute_opacity at N > 1) with uneven row counts
    and three overlapped chunks, on CPU tensors over gloo."""
    import torch.multiprocessing as tmp_mp
    port = 31500 + os.getpid() % 2000
    tmp_mp.spawn(_gloo_assembler_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    full = np.arange(23 * 17, dtype=np.double).reshape(23, 17) ** 1.5
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"asm_{r}.npy"), full)


def test_equal_count_partition():
    from pyratbay_b200 import parallel
    cost = np.random.default_rng(1).uniform(1, 30, 1020)
    owners = parallel.unit_owners(1020, 8, cost, equal_counts=True)
    assert sorted(np.concatenate(owners)) == list(range(1020))
    assert max(len(o) for o in owners) == 128 and min(len(o) for o in owners) >= 124
    loads = np.array([cost[o].sum() for o in owners])
    assert loads.max() / loads.min() < 1.02
    assert all(np.all(np.diff(o) > 0) for o in owners)


def test_partition_by_temperature():
    """The product partition of a table: contiguous pieces of a snake-ordered temperature
    sequence -- every unit exactly once, balanced cost, few temperatures per rank."""
    from pyratbay_b200 import parallel
    ntemp, nlayers = 20, 51
    p_cost = 28.0 + 32.0 * np.linspace(0, 1, nlayers) ** 2           # grows with pressure
    cost = np.concatenate([p_cost * (1.0 + 0.02 * t) for t in range(ntemp)])   # and with T
    for world in (2, 3, 4, 8):
        owners = parallel.unit_owners_by_temperature(ntemp, nlayers, world, cost)
        assert len(owners) == world
        assert sorted(np.concatenate(owners)) == list(range(ntemp * nlayers))
        loads = np.array([cost[o].sum() for o in owners])
        assert loads.max() / loads.min() < 1.06
        temps_per_rank = [len(np.unique(o // nlayers)) for o in owners]
        assert max(temps_per_rank) <= -(-ntemp // world) + 2
        assert all(np.all(np.diff(o) > 0) for o in owners)
    assert len(parallel.unit_owners_by_temperature(ntemp, nlayers, 1)[0]) == ntemp * nlayers


def test_unit_costs_balance_table_partition():
    """Cost model used to deal (T,p) units to ranks: grows with pressure (wider profiles) and
    gives a far better balance than round-robin would on cost."""
    from pyratbay_b200 import parallel
    c = helpers.mock_case(with_profile=False)
    ntemp, nlayers = 10, c.atm.nlayers
    temps = np.repeat(np.linspace(300, 3000, ntemp), nlayers)
    press = np.tile(c.atm.press, ntemp)
    vmr = np.tile(c.atm.vmr, (ntemp, 1))
    cost = parallel.unit_costs(c.voigt, c.spec, c.atm, c.iso_atm_index, c.iso_mass, temps,
                               press, vmr)
    assert cost.shape == (ntemp * nlayers,) and np.all(cost > 0)
    per_layer = cost[:nlayers]
    assert per_layer[-1] > per_layer[0]          # 100 bar costs more than 1e-6 bar
    parts = [parallel.partition_units(len(cost), r, 8, cost) for r in range(8)]
    assert sorted(np.concatenate(parts)) == list(range(len(cost)))
    loads = np.array([cost[p].sum() for p in parts])
    assert loads.max() / loads.min() < 1.03


def test_str_of_voigt_and_line_by_line_match_reference():
    """`str()` of the reference-shaped objects: the Voigt text pinned by the reference's
    tests/test_str.py:338-366 (golden string + grid from the reference run; the profile values
    are the golden edge profiles, no GPU needed) and the Line_By_Line text for the mock TLI."""
    from types import SimpleNamespace
    from pyratbay_b200.spectrum import Spectrum
    from pyratbay_b200.voigt import Voigt
    from pyratbay_b200.line_by_line import Line_By_Line
    g = helpers.golden("voigt_h2o_1.1-1.7um.npz")
    spec = Spectrum(wl_low=1.1 * pc.um, wl_high=1.7 * pc.um, wnstep=1.0, wnosamp=2160)
    atm = helpers.mock_atmosphere()
    v = Voigt(spec, atm, np.full(4, 5), None, extent=100.0, cutoff=25.0)
    v.size, v.index = g["size"], g["index"]
    prof = np.zeros(int(g["profile_len"]))
    for key, (m, n) in {"profile_first": (0, 0), "profile_last": (-1, -1)}.items():
        prof[v.index[m, n]:v.index[m, n] + len(g[key])] = g[key]
    v._profile = prof
    assert str(v) == str(g["voigt_str"])

    case = helpers.mock_case(with_profile=False)
    fake = SimpleNamespace(spec=case.spec, atm=case.atm)
    lbl = Line_By_Line(case.tlifile, case.atm.species, case.spec.wnlow, case.spec.wnhigh, fake)
    want = str(helpers.golden("mock_forward.npz")["lbl_str"])
    assert str(lbl).replace(str(lbl.tlifile), "['TLI']") == want


def test_nearest_thresholds_reproduce_the_nearest_sample_search():
    """pb200_nearest_thresholds (host): counting thresholds <= v equals the reference's
    nearest-sample search with ties to the lower index (utils.h:44-72, 75-89)."""
    from pyratbay_b200.engine import nearest_thresholds
    from pyratbay_b200._lib import PB200Error
    rng = np.random.default_rng(3)
    for grid in (np.logspace(-3, 0, 40), np.array([1.0, 2.0, 2.5, 7.0]), np.linspace(0.1, 5, 17)):
        thr = nearest_thresholds(grid)
        assert thr[0] == 0.0 and np.all(np.diff(thr) > 0)
        v = np.concatenate([rng.uniform(0, grid[-1] * 1.3, 20000), grid, thr[1:],
                            np.nextafter(thr[1:], 0), 0.5 * (grid[1:] + grid[:-1])])
        got = np.searchsorted(thr, v, side="right") - 1
        hi = np.clip(np.searchsorted(grid, v, side="left"), 1, len(grid) - 1)
        lo = hi - 1
        want = np.where(np.abs(grid[hi] - v) < np.abs(grid[lo] - v), hi, lo)
        want = np.where(v < grid[0], 0, np.where(v > grid[-1], len(grid) - 1, want))
        assert np.array_equal(got, want)
    with pytest.raises(PB200Error):
        nearest_thresholds(np.array([1.0, 1.0, 2.0]))


def test_shim_static_keys_and_install(tmp_path):
    """Drop-in shim (pyratbay_b200/shim), host side only: content keys of static arguments
    (small arrays re-hashed on every call so an in-place change is seen, large ones cached per
    array object) and the two forwarding modules written into a reference lib/ directory."""
    from pyratbay_b200.shim import client, install_into
    small = np.arange(1000, dtype=np.float64)
    k1 = client.static_key(small)
    small[10] += 1.0
    assert client.static_key(small) != k1
    assert client.static_key(small.copy()) == client.static_key(small)     # content, not identity
    big = np.zeros(200_000, np.float64)                                     # 1.6 MB: cached
    kb = client.static_key(big)
    assert client.static_key(big) == kb and kb != client.static_key(np.ones(200_000))
    assert client.static_key(np.zeros(200_000, np.float64)) == kb
    assert client.static_key(np.arange(5, dtype=np.int64)) != client.static_key(np.arange(5.0))
    lib = tmp_path / "lib"
    lib.mkdir()
    (lib / "_extcoeff.cpython-312-x86_64-linux-gnu.so").write_bytes(b"x")
    (lib / "_trapezoid.cpython-312-x86_64-linux-gnu.so").write_bytes(b"x")
    install_into(str(lib))
    names = sorted(os.listdir(lib))
    assert names == ["_extcoeff.py", "_trapezoid.cpython-312-x86_64-linux-gnu.so", "vprofile.py"]
    assert "pyratbay_b200.shim._extcoeff" in (lib / "_extcoeff.py").read_text()
    import inspect
    from pyratbay_b200.shim import _extcoeff, vprofile
    assert len(inspect.signature(_extcoeff.extinction).parameters) == 27    # _extcoeff.c:114-123
    assert list(inspect.signature(vprofile.grid).parameters) == [
        "profile", "psize", "index", "lorentz", "doppler", "dwn", "verb"]  # vprofile.c:53-57


def test_streaming_opacity_writer_matches_write_opacity(tmp_path):
    """io.OpacityWriter (rows appended while later ones are still being computed) writes the
    members np.savez writes in io.write_opacity (io/io.py:570-606): same arrays, dtypes, units."""
    from pyratbay_b200 import io
    rng = np.random.default_rng(4)
    temp, press, wn = np.linspace(300, 3000, 4), np.logspace(-6, 2, 5), np.linspace(1000, 1100, 7)
    table = rng.random((4, 5, 7))
    a, b = str(tmp_path / "a.npz"), str(tmp_path / "b")      # .npz appended like np.savez does
    io.write_opacity(a, "H2O", temp, press, wn, table)
    with io.OpacityWriter(b, "H2O", temp, press, wn) as w:
        flat = table.reshape(20, 7)
        w.write(flat[:3])
        w.write(flat[3:11])
        w.write(flat[11:])
    fa, fb = np.load(a, allow_pickle=True), np.load(b + ".npz", allow_pickle=True)
    assert sorted(fa.files) == sorted(fb.files)
    for key in fa.files:
        assert fa[key].dtype == fb[key].dtype and fa[key].shape == fb[key].shape
        if key == "units":
            assert fa[key].item() == fb[key].item()
        else:
            assert np.array_equal(fa[key], fb[key])
    assert all(np.array_equal(x, y) if isinstance(x, np.ndarray) else x == y
               for x, y in zip(io.read_opacity(a), io.read_opacity(b + ".npz")))
    with pytest.raises(ValueError):                          # incomplete table
        with io.OpacityWriter(str(tmp_path / "c.npz"), "H2O", temp, press, wn) as w:
            w.write(flat[:5])
