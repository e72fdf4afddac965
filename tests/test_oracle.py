"""CPU tests that pin the oracle (oracle/lbl_oracle.c) against the reference:
golden vectors produced by the unmodified reference (tests/golden/make_golden.py), the
reference's own Voigt known answers (tests/test_str.py:338-366) and, when oracle/_ref has
been built, the reference's compiled modules on seeded inputs."""
import os

import numpy as np
import pytest

import helpers
from pyratbay_b200 import constants as pc

orc = helpers.oracle_module()


def test_oracle_voigt_matches_reference_table():
    case = helpers.mock_case()
    g = helpers.golden("mock_voigt.npz")
    assert np.array_equal(case.size, g["size"]) and np.array_equal(case.index, g["index"])
    assert len(case.profile) == int(g["profile_len"])
    want, got = g["profile_strided"], case.profile[::int(g["stride"])]
    nz = want > 0
    assert np.max(np.abs(got[nz] / want[nz] - 1)) < 1e-13
    assert np.all(got[~nz] == 0)
    for key, (m, n) in {"profile_first": (0, 0), "profile_mid": (40, 10),
                        "profile_last": (-1, -1)}.items():
        i0, half = case.index[m, n], case.size[m, n]
        assert np.max(np.abs(case.profile[i0:i0 + 2 * half + 1] / g[key] - 1)) < 1e-13
    # zero padding at the end of the table (skipped profiles count one sample, voigt.py:142)
    tail, want_tail = case.profile[-2000:], g["profile_tail"]
    assert np.array_equal(tail == 0, want_tail == 0) and np.sum(want_tail == 0) > 100
    np.testing.assert_allclose(tail, want_tail, rtol=1e-13)


def test_oracle_voigt_known_answers_of_reference_test_str():
    g = helpers.golden("voigt_h2o_1.1-1.7um.npz")
    text = str(g["voigt_str"])
    # literal lines of the reference's tests/test_str.py:338-366
    for line in ["[[ 1072  1120 ...  8687  9074]", " [54000 54000 ... 54000 54000]]",
                 "[[       0     2145 ...   341896   359271]",
                 " [47041097 47041097 ... 47041097 47041097]]",
                 "profile[ 0, 0]: [2.85914e-08 2.86448e-08 ... 2.86448e-08 2.85914e-08]",
                 "profile[99,49]: [4.99389e-03 4.99404e-03 ... 4.99404e-03 4.99389e-03]"]:
        assert line in text
    # the oracle reproduces two whole profiles of that grid; [99,49] is an alias of [99,0]
    # (Doppler/Lorentz ratio below dlratio, vprofile.c:99-105)
    assert g["index"][-1, -1] == g["index"][-1, 0]
    for key, (m, n) in {"profile_first": (0, 0), "profile_last": (-1, 0)}.items():
        half = int(g["size"][m, n])
        size = np.array([[half]], np.int64)
        index = np.zeros((1, 1), np.int64)
        prof = np.zeros(2 * half + 1)
        orc.grid(prof, size, index, g["lorentz"][[m]], g["doppler"][[n]], float(g["ownstep"]))
        assert np.max(np.abs(prof / g[key] - 1)) < 1e-13


@pytest.mark.skipif(not os.path.exists("/root/reference/src_c/include/voigt.h"),
                    reason="reference tree not present")
def test_series_coefficients_equal_reference_literals():
    import re
    text = open("/root/reference/src_c/include/voigt.h").read()
    block = text[text.index("static double ferf"):text.index("int _voigt_maxelements=")]
    body = block[block.index("{") + 1:]
    lits = [float(x) for x in re.findall(r"([0-9]\.[0-9]+(?:e-?[0-9]+)?)\s*,?\s*//", body)]
    lits.append(float(re.findall(r"([0-9]\.[0-9]+e-?[0-9]+)\s*\}", body)[0]))
    assert len(lits) == 61
    # same extended-precision recurrence as oracle/lbl_oracle.c and csrc/voigt.cu
    fact, mine = np.longdouble(1), []
    for n in range(61):
        if n > 0:
            fact = fact * np.longdouble(n)
        mine.append(float(np.longdouble(1) / (fact * np.longdouble(2 * n + 1))))
    assert mine == lits   # bit-identical doubles


@pytest.mark.parametrize("resolution,golden_file", [
    (None, "mock_opacity_table.npz"), (15000.0, "mock_opacity_table_R.npz")])
def test_oracle_extinction_matches_reference_table(resolution, golden_file):
    case = helpers.mock_case(resolution=resolution)
    g = helpers.golden(golden_file)
    z = helpers.partition(case, g["temp"])
    assert np.array_equal(g["wn"], case.spec.wn)
    worst = 0.0
    for it in range(0, 10, 3):
        for il in range(0, 51, 7):
            temp = g["temp"][it]
            dens = case.atm.vmr[il] * case.atm.press[il] * pc.bar / (pc.k * temp)
            ext = np.zeros((1, case.spec.nwave))
            orc.extinction(ext, *case.unit_args(temp, dens, z[:, it]), 0, 0,
                           int(resolution is not None))
            want = g["etable"][it, il]
            worst = max(worst, np.max(np.abs(ext[0] - want)) / np.max(want))
    assert worst < 1e-13


def test_oracle_forward_model_matches_reference():
    case = helpers.mock_case()
    g = helpers.golden("mock_forward.npz")
    z = helpers.partition(case, g["temp"])
    for il in (0, 17, 31, 50):
        ext = np.zeros((1, case.spec.nwave))
        orc.extinction(ext, *case.unit_args(g["temp"][il], g["d"][il], z[:, il]), 0, 1, 0)
        assert np.max(np.abs(ext[0] - g["ec_all"][il])) / np.max(g["ec_all"][il]) < 1e-13
    ext = np.zeros((1, case.spec.nwave))
    orc.extinction(ext, *case.unit_args(g["temp"][31], g["d"][31], z[:, 31]), 0, 0, 0)
    want = g["ec_layer31"][0]
    got = ext[0] * g["d"][31, 5]          # H2O is species 5 of the test atmosphere
    assert np.max(np.abs(got - want)) / np.max(want) < 1e-13


@pytest.mark.parametrize("kwargs", [dict(), dict(ethresh=1e-6), dict(resolution=8000.0),
                                    dict(cutoff=0.0, extent=20.0)])
def test_oracle_matches_compiled_reference_on_seeded_inputs(kwargs):
    ec_ref, vp_ref = orc.load_ref()
    if ec_ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    case = helpers.synthetic_case(nlines=6000, **kwargs)
    # Voigt grid: reference module vs oracle
    size, index = case.size_in.copy(), np.zeros_like(case.size_in)
    prof = np.zeros(case.voigt.profile_len)
    vp_ref.grid(prof, size, index, case.lorentz, case.doppler, case.spec.ownstep, 0)
    assert np.array_equal(size, case.size) and np.array_equal(index, case.index)
    nz = prof > 0
    assert np.max(np.abs(case.profile[nz] / prof[nz] - 1)) < 1e-13
    temps, dens = case.atm.temp, case.atm.d
    isoz = helpers.partition(case, temps).T
    res = int(case.spec.interpolate)
    for add in (0, 1):
        for u in range(0, len(temps), 2):
            a = np.zeros((1, case.spec.nwave))
            b = np.zeros((1, case.spec.nwave))
            args = case.unit_args(temps[u], dens[u], isoz[u])
            ec_ref.extinction(a, *args, 0, add, res)
            orc.extinction(b, *args, 0, add, res)
            assert np.max(np.abs(a - b)) / np.max(a) < 1e-13


def test_oracle_interp_ec_matches_reference_line_sample():
    import pyratbay_b200 as pb
    g = helpers.golden("mock_line_sample.npz")
    _, _, temp, press, wn, table = pb.io.read_opacity(
        os.path.join(helpers.GOLDEN, "mock_opacity_file.npz"))
    table = table[np.newaxis]
    nlayers, nwave = len(press), len(wn)
    ext = np.zeros((nlayers, nwave))
    orc.interp_ec(ext, table, temp, g["temperature"], g["density"], 0, nlayers)
    np.testing.assert_allclose(ext, g["ec"], rtol=1e-14)
    cs = np.zeros((1, nlayers, nwave))
    orc.interp_ec_per_mol(cs, table, temp, g["temperature"], np.ones((nlayers, 1)), 0, nlayers)
    np.testing.assert_allclose(cs, g["cs_per_mol"], rtol=1e-14)


def _od_args(g, name):
    kw = {"emission": dict(transit=False, maxdepth=np.inf, itop=0, ibottom=None),
          "emission_max": dict(transit=False, maxdepth=10.0, itop=3, ibottom=45),
          "transit": dict(transit=True, maxdepth=np.inf, itop=0, ibottom=None),
          "transit_max": dict(transit=True, maxdepth=10.0, itop=2, ibottom=48)}[name]
    return kw


@pytest.mark.parametrize("name", ["emission", "emission_max", "transit", "transit_max"])
def test_oracle_optical_depth_matches_reference(name):
    """Optical depth (next-tier row) against the reference's optical_depth on the golden
    forward-model extinction (tests/golden/make_golden.py section 4b)."""
    from pyratbay_b200.optic_depth import transit_path, _path_matrix
    g = helpers.golden("mock_optical_depth.npz")
    ec = np.ascontiguousarray(g["ec"] * 3e4)
    nlayers, nwave = ec.shape
    kw = _od_args(g, name)
    ibottom = nlayers if kw["ibottom"] is None else kw["ibottom"]
    depth = np.zeros((nlayers, nwave))
    ideep = np.zeros(nwave, np.int32)
    if kw["transit"]:
        paths = _path_matrix(transit_path(g["radius"], kw["itop"]), nlayers)
        orc.transit_optical_depth(depth, ideep, ec, paths, kw["maxdepth"], kw["itop"], ibottom)
    else:
        orc.plane_parallel_optical_depth(depth, ideep, ec, -np.ediff1d(g["radius"]),
                                         kw["maxdepth"], kw["itop"], ibottom)
    np.testing.assert_allclose(depth, g[name + "_depth"], rtol=1e-13, atol=0)
    assert np.array_equal(ideep, g[name + "_ideep"])
