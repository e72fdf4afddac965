"""Worker of tests/test_gpu_dropin.py.  Runs with PYTHONPATH = oracle/refstubs :
oracle/_ref/shimmed : repo root, i.e. `import pyratbay` is the UNMODIFIED reference package whose
lib/_extcoeff and lib/vprofile forward to the GPU engine (pyratbay_b200/shim).  It runs the
reference's own drivers (pb.run on a tli and an opacity config, Line_By_Line and Line_Sample
calls) exactly as tests/golden/make_golden.py did with the compiled C modules, and stores the
outputs for comparison with those goldens."""
import os
import shutil
import sys
import time

import numpy as np


def main():
    golden, run = sys.argv[1], sys.argv[2]
    import pyratbay as pb
    import pyratbay.io as io
    import pyratbay.opacity as op
    from pyratbay.lib import _extcoeff as ec, vprofile as vp
    assert ec.__name__.startswith("pyratbay_b200.shim") or \
        ec.extinction.__module__.startswith("pyratbay_b200.shim")
    assert vp.grid.__module__.startswith("pyratbay_b200.shim")
    os.makedirs(os.path.join(run, "outputs"), exist_ok=True)
    os.makedirs(os.path.join(run, "inputs"), exist_ok=True)
    shutil.copy(os.path.join(golden, "inputs", "Mock_HITRAN_H2O_1.00-1.01um.par"),
                os.path.join(run, "inputs"))
    shutil.copy(os.path.join(golden, "inputs", "atmosphere_uniform_test.atm"),
                os.path.join(run, "inputs"))
    os.chdir(run)
    with open("tli.cfg", "w") as f:
        f.write("[pyrat]\nrunmode = tli\nlogfile = outputs/mock.log\n"
                "dblist = inputs/Mock_HITRAN_H2O_1.00-1.01um.par\ndbtype = hitran\npflist = tips\n"
                "wl_low  = 1.00 um\nwl_high = 1.01 um\nverb = 1\n")
    pb.run("tli.cfg")
    base = ("[pyrat]\nrunmode = opacity\natmfile = inputs/atmosphere_uniform_test.atm\n"
            "tlifile = outputs/mock.tli\nwl_low   = 1.00 um\nwl_high  = 1.01 um\nwnosamp = 2160\n"
            "voigt_extent = 100.0\ntmin  =  300\ntmax  = 3000\ntstep =  300\nncpu = 3\nverb = 1\n")
    with open("opacity.cfg", "w") as f:
        f.write(base + "logfile = outputs/table.log\nwnstep = 1.0\n")
    t0 = time.time()
    pyrat = pb.run("opacity.cfg")                     # forks 3 workers -> shim -> GPU server
    t_table = time.time() - t0
    ex, atm = pyrat.ex, pyrat.atm
    out = {"etable": ex.etable, "file_etable": io.read_opacity(ex.sampled_cs[0], extract="opacity"),
           "table_s": t_table, "n_units": ex.ntemp * ex.nlayers}
    v = pyrat.voigt
    out["voigt_size"], out["voigt_index"] = v.size, v.index
    out["profile_strided"] = v.profile[::997]
    out["profile_sum"] = np.sum(v.profile)

    # constant-R table (2-point interpolation path)
    with open("opacity_R.cfg", "w") as f:
        f.write(base + "logfile = outputs/table_R.log\nresolution = 15000.0\nwnstep = 1.0\n")
    out["etable_R"] = pb.run("opacity_R.cfg").ex.etable

    # forward model through the reference's Line_By_Line (forked workers, add=1) and get_ec
    lbl = pyrat.opacity.models[pyrat.opacity.models_type.index('lbl')]
    dens = atm.d[:, lbl.mol_index]
    t0 = time.time()
    out["ec_all"] = np.copy(lbl.calc_extinction_coefficient(atm.temp, dens))
    out["forward_s"] = time.time() - t0
    out["ec_layer31"] = np.copy(lbl.calc_extinction_coefficient(atm.temp, dens, layer=31))
    out["ec_skip"] = np.copy(lbl.calc_extinction_coefficient(atm.temp, dens, skip_mol=['H2O']))

    # the reference's Line_Sample on the table it just wrote (interp_ec through the shim)
    ls = op.Line_Sample(ex.sampled_cs[0])
    temp = np.linspace(450.0, 2900.0, ls.nlayers)
    out["ls_cs"] = ls.calc_cross_section(temp)
    out["ls_cs_per_mol"] = ls.calc_cross_section(temp, per_mol=True)
    out["ls_ec"] = ls.calc_extinction_coefficient(temp, dens)
    out["ls_ec_layer"] = ls.calc_extinction_coefficient(temp, dens, layer=20)

    # per-call overhead of the 27-argument call once everything is resident
    from pyratbay.pyrat import extinction as ref_ex
    t0 = time.time()
    for _ in range(20):
        ref_ex.extinction(pyrat, [25], grid=False, add=False)
    out["per_call_s"] = (time.time() - t0) / 20
    from pyratbay_b200.shim import client
    out["server_stats"] = np.array(str(client.request("stats")))
    np.savez(os.path.join(run, "dropin_results.npz"), **out)
    print("DROPIN OK", flush=True)


if __name__ == "__main__":
    main()
