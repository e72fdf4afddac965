#!/usr/bin/env python
"""Golden fixtures for the line-list readers / TLI writer, made by
running the UNMODIFIED reference (`pbay -c tli`) on its own mock inputs.

Runs only in the build container (needs /root/reference); same scratch copy + stand-in
modules as make_golden.py.  Writes:
  tests/golden/inputs/*            the reference's small mock line lists and partition files
                                   (data fixtures of its tests/inputs, needed as INPUT here)
  tests/golden/exomol_nh3.tli      reference TLI of tests/configs/tli_exomol_test.cfg
  tests/golden/repack_co2.tli      reference TLI of tests/configs/tli_repack_test.cfg
  tests/golden/hitran_h2o_pf.tli   reference TLI of the mock HITRAN H2O list with a tabulated
                                   partition file (written here from the TIPS table)
  pyratbay_b200/data/isotopes_subset.json
                                   isotope names / ratios / masses and HITRAN order of the
                                   molecules of the benchmark configs (from the reference's data)
  pyratbay_b200/data/tips_subset.npz
                                   TIPS-2021 partition functions of the same molecules
  tests/golden/repack_co2_tips.tli reference TLI of the repack CO2 list with pflist = tips
(mock_hitran_h2o.tli with pflist=tips comes from make_golden.py.)
"""
import json
import os
import shutil

import numpy as np

from make_golden import REF, HERE, REPO, WORK, build_scratch, write_cfg

INPUTS = [
    "Mock_HITRAN_H2O_1.00-1.01um.par",
    "14N-1H3__MockBYTe__04999-05000.trans", "14N-1H3__MockBYTe.states",
    "15N-1H3__MockBYTe-15__04999-05000.trans", "15N-1H3__MockBYTe-15.states",
    "PF_Exomol_NH3.dat", "PF_tips_CO2.dat",
    "CO2_hitran_2.50-2.52um_repack-0.01_lbl.dat",
]
MOLECULES = {"H2O": 1, "CO2": 2, "CO": 5, "CH4": 6, "NH3": 11, "HCN": 23}


def main():
    build_scratch()
    import pyratbay as pb
    import pyratbay.io as io
    import pyratbay.constants as pc
    import pyratbay.opacity.partitions as pf

    run = os.path.join(WORK, "run_tli")
    os.makedirs(os.path.join(run, "outputs"), exist_ok=True)
    os.makedirs(os.path.join(run, "inputs"), exist_ok=True)
    gin = os.path.join(HERE, "inputs")
    os.makedirs(gin, exist_ok=True)
    for f in INPUTS:
        shutil.copy(os.path.join(REF, "tests", "inputs", f), os.path.join(run, "inputs", f))
        shutil.copy(os.path.join(REF, "tests", "inputs", f), os.path.join(gin, f))
        os.chmod(os.path.join(gin, f), 0o644)
    os.chdir(run)

    write_cfg("exomol.cfg", """runmode = tli
logfile = outputs/exomol_nh3.log
dblist =
    inputs/14N-1H3__MockBYTe__04999-05000.trans
    inputs/15N-1H3__MockBYTe-15__04999-05000.trans
dbtype = exomol exomol
pflist = inputs/PF_Exomol_NH3.dat
wl_low  = 2.0 um
wl_high = 2.00002 um
verb = 1
""")
    pb.run("exomol.cfg")
    shutil.copy("outputs/exomol_nh3.tli", os.path.join(HERE, "exomol_nh3.tli"))

    write_cfg("repack.cfg", """runmode = tli
logfile = outputs/repack_co2.log
dblist = inputs/CO2_hitran_2.50-2.52um_repack-0.01_lbl.dat
dbtype = repack
pflist = inputs/PF_tips_CO2.dat
wl_low  = 2.50 um
wl_high = 2.52 um
verb = 1
""")
    pb.run("repack.cfg")
    shutil.copy("outputs/repack_co2.tli", os.path.join(HERE, "repack_co2.tli"))

    # HITRAN with a tabulated partition file (the file itself is a fixture, too)
    pf.tips("H2O", outfile="inputs/PF_tips_H2O.dat")
    shutil.copy("inputs/PF_tips_H2O.dat", os.path.join(gin, "PF_tips_H2O.dat"))
    write_cfg("hitran_pf.cfg", """runmode = tli
logfile = outputs/hitran_h2o_pf.log
dblist = inputs/Mock_HITRAN_H2O_1.00-1.01um.par
dbtype = hitran
pflist = inputs/PF_tips_H2O.dat
wl_low  = 1.002 um
wl_high = 1.008 um
verb = 1
""")
    pb.run("hitran_pf.cfg")
    shutil.copy("outputs/hitran_h2o_pf.tli", os.path.join(HERE, "hitran_h2o_pf.tli"))

    # Isotope data of the benchmark molecules
    mol, hit_iso, exo_iso, ratio, mass = io.read_isotopes(pc.ROOT + "pyratbay/data/isotopes.dat")
    data = {"hitran_mol_id": {str(v): k for k, v in MOLECULES.items()}, "molecules": {}}
    for name in MOLECULES:
        sel = [i for i in range(len(mol)) if mol[i] == name]
        data["molecules"][name] = {
            "hitran_iso": [str(hit_iso[i]) for i in sel],
            "exomol_iso": [str(exo_iso[i]) for i in sel],
            "ratio": [float(ratio[i]) for i in sel],
            "mass": [float(mass[i]) for i in sel],
            # HITRAN isotope order as the reference's Hitran reader takes it from TIPS
            "tips_order": [str(s) for s in pf.tips(name)[1]],
        }
    with open(os.path.join(REPO, "pyratbay_b200", "data", "isotopes_subset.json"), "w") as f:
        json.dump(data, f, indent=1)

    # TIPS-2021 partition functions of the same molecules (pflist = tips)
    tables = {}
    for name in MOLECULES:
        z, iso, temp = pf.tips(name)
        tables[f"{name}_temp"] = np.asarray(temp, np.double)
        tables[f"{name}_z"] = np.asarray(z, np.double)
        tables[f"{name}_iso"] = np.array([str(i) for i in iso])
    np.savez_compressed(os.path.join(REPO, "pyratbay_b200", "data", "tips_subset.npz"), **tables)

    # HITRAN CO2 with pflist = tips through the reference (golden for the bundled table)
    write_cfg("repack_tips.cfg", """runmode = tli
logfile = outputs/repack_co2_tips.log
dblist = inputs/CO2_hitran_2.50-2.52um_repack-0.01_lbl.dat
dbtype = repack
pflist = tips
wl_low  = 2.50 um
wl_high = 2.52 um
verb = 1
""")
    pb.run("repack_tips.cfg")
    shutil.copy("outputs/repack_co2_tips.tli", os.path.join(HERE, "repack_co2_tips.tli"))
    print("done")


if __name__ == "__main__":
    main()
