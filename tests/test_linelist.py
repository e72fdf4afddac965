"""Line-list readers and make_tli against TLI files written by the
UNMODIFIED reference from its own mock inputs (tests/golden/make_golden_tli.py,
make_golden.py): the output files must be byte-identical.  CPU only."""
import os

import numpy as np
import pytest

from pyratbay_b200 import linelist, lread, tli as ptli

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
INP = os.path.join(GOLD, "inputs")


def _same_file(a, b):
    with open(a, "rb") as fa, open(b, "rb") as fb:
        return fa.read() == fb.read()


def test_hitran_tips_tli_is_byte_identical(tmp_path):
    # reference: tests/configs/tli_hitran_test.cfg retargeted to 1.00-1.01 um (make_golden.py)
    out = str(tmp_path / "h2o.tli")
    lread.make_tli(os.path.join(INP, "Mock_HITRAN_H2O_1.00-1.01um.par"), "tips", "hitran",
                   out, 1.00, 1.01, "um")
    assert _same_file(out, os.path.join(GOLD, "mock_hitran_h2o.tli"))


def test_hitran_partition_file_and_narrow_window(tmp_path):
    out = str(tmp_path / "h2o_pf.tli")
    lread.make_tli([os.path.join(INP, "Mock_HITRAN_H2O_1.00-1.01um.par")],
                   [os.path.join(INP, "PF_tips_H2O.dat")], ["hitran"], out, 1.002, 1.008, "um")
    assert _same_file(out, os.path.join(GOLD, "hitran_h2o_pf.tli"))


def test_exomol_two_isotopologues_one_database(tmp_path):
    # reference: tests/configs/tli_exomol_test.cfg, known answers of tests/test_tli.py:116-136
    out = str(tmp_path / "nh3.tli")
    dbs = lread.make_tli(
        [os.path.join(INP, "14N-1H3__MockBYTe__04999-05000.trans"),
         os.path.join(INP, "15N-1H3__MockBYTe-15__04999-05000.trans")],
        os.path.join(INP, "PF_Exomol_NH3.dat"), ["exomol", "exomol"], out, 2.0, 2.00002, "um")
    assert _same_file(out, os.path.join(GOLD, "exomol_nh3.tli"))
    assert len(dbs) == 1 and dbs[0].name == "Exomol NH3" and dbs[0].ntemp == 2000
    assert list(dbs[0].iso_name) == ["4111", "5111"]
    _, wn, gf, elow, iso = ptli.read_tli_file(out, 0.0, 1e5)
    assert len(wn) == 1000 and np.bincount(iso).tolist() == [500, 500]


def test_repack_tli_is_byte_identical(tmp_path):
    # reference: tests/configs/tli_repack_test.cfg, known answers of tests/test_tli.py:140-157
    out = str(tmp_path / "co2.tli")
    dbs = lread.make_tli(os.path.join(INP, "CO2_hitran_2.50-2.52um_repack-0.01_lbl.dat"),
                         os.path.join(INP, "PF_tips_CO2.dat"), "repack", out, 2.50, 2.52, "um")
    assert _same_file(out, os.path.join(GOLD, "repack_co2.tli"))
    assert dbs[0].name == "repack hitran CO2" and dbs[0].ntemp == 1001
    assert list(dbs[0].iso_name) == ["266", "366", "628", "627"]


def test_missing_partition_isotopes_and_bad_inputs(tmp_path):
    # tests/test_tli.py:160-167: partition file without the line list's isotopes
    pf_bad = tmp_path / "pf.dat"
    pf_bad.write_text("@ISOTOPES\n 9999\n@DATA\n 100.0 1.0\n 200.0 2.0\n")
    with pytest.raises(ValueError, match="No partition functions found for these isotopes"):
        lread.make_tli(os.path.join(INP, "14N-1H3__MockBYTe__04999-05000.trans"), str(pf_bad),
                       "exomol", str(tmp_path / "x.tli"), 2.0, 2.00002, "um")
    with pytest.raises(ValueError, match="Unknown type"):
        lread.make_tli("a.par", "tips", "nosuchdb", str(tmp_path / "x.tli"), 1.0, 2.0, "um")
    with pytest.raises(ValueError, match="does not match"):
        lread.make_tli(["a", "b"], ["p", "q", "r"], ["hitran"], str(tmp_path / "x.tli"),
                       1.0, 2.0, "um")


def test_window_search_matches_reference_semantics():
    """driver.py:80-137: binary search, then a linear walk.  Known answers produced by the
    reference's Linelist.binsearch on the same array (note that searching DOWN for a repeated
    value stops at the record just above the run: the reference's own behaviour)."""
    wn = np.array([1.0, 2.0, 2.0, 2.0, 3.0, 4.0, 4.0, 5.0])
    f = linelist.Linelist.binsearch_array
    get = wn.__getitem__
    want = {(2.0, False): 4, (2.0, True): 3, (0.5, False): 0, (9.0, True): 7,
            (3.5, False): 5, (3.5, True): 4, (4.0, False): 7, (4.0, True): 6}
    for (target, up), irec in want.items():
        assert f(get, target, 0, 7, up) == irec


def test_exomol_file_name_parser():
    # doctest values of tools/tools.py:864-881
    cases = {
        '1H2-16O__POKAZATEL__00400-00500.trans.bz2': ('H2O', '116'),
        '1H-2H-16O__VTT__00250-00500.trans.bz2': ('H2O', '126'),
        '12C-16O2__HITEMP.pf': ('CO2', '266'),
        '12C-16O-18O__Zak.par': ('CO2', '268'),
        '12C-1H4__YT10to10__01100-01200.trans.bz2': ('CH4', '21111'),
        '12C-1H3-2H__MockName__01100-01200.trans.bz2': ('CH4', '21112'),
    }
    for name, want in cases.items():
        assert linelist.get_exomol_mol(name) == want


def test_runmode_tli_through_the_driver(tmp_path):
    """`pbay -c tli.cfg` (driver.py:35-46): the reference's tli_exomol_test.cfg keys."""
    from pyratbay_b200 import pyrat as pb
    cfg = tmp_path / "tli.cfg"
    cfg.write_text(f"""[pyrat]
runmode = tli
logfile = {tmp_path}/ExoMol_NH3.log
dblist =
    {INP}/14N-1H3__MockBYTe__04999-05000.trans
    {INP}/15N-1H3__MockBYTe-15__04999-05000.trans
dbtype = exomol exomol
pflist = {INP}/PF_Exomol_NH3.dat
wl_low  = 2.0 um
wl_high = 2.00002 um
verb = 0
""")
    assert pb.run(str(cfg)) is None
    assert _same_file(str(tmp_path / "ExoMol_NH3.tli"), os.path.join(GOLD, "exomol_nh3.tli"))


def test_bundled_tips_tables(tmp_path):
    """pflist = tips for a molecule other than H2O: TLI byte-identical to the reference's."""
    out = str(tmp_path / "co2_tips.tli")
    lread.make_tli(os.path.join(INP, "CO2_hitran_2.50-2.52um_repack-0.01_lbl.dat"), "tips",
                   "repack", out, 2.50, 2.52, "um")
    assert _same_file(out, os.path.join(GOLD, "repack_co2_tips.tli"))
