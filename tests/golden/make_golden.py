#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  It makes a scratch copy of the
reference package under /tmp, compiles its C modules with the reference's flags
(setup.py:19), imports it with the stand-in modules of oracle/refstubs for the absent
third-party packages (mc3, matplotlib, chemcat, h5py, cycler) and stores small
input/output vectors.  Nothing from the reference's sources is written into the repo.

Fixtures written (all small):
  mock_hitran_h2o.tli          TLI built by the reference from its own
                               tests/inputs/Mock_HITRAN_H2O_1.00-1.01um.par (888 lines)
  mock_atmosphere.npz          the reference's tests/inputs/atmosphere_uniform_test.atm as arrays
  mock_opacity_table.npz       pb.run(opacity cfg): etable[10,51,100] + grids  (resample mode)
  mock_opacity_table_R.npz     same with resolution=15000 (constant-R, linterp mode)
  mock_forward.npz             Line_By_Line.calc_extinction_coefficient (add=1) + get_ec(layer)
  mock_voigt.npz               Voigt sizes/indices/grids + sampled profile values, 1.00-1.01 um
  voigt_h2o_1.1-1.7um.npz      Voigt grid of the reference's test_str.py:338-366 case
  mock_line_sample.npz         op.Line_Sample on the table: interp_ec / interp_ec_per_mol outputs
  tli_window_cases.npz         read_tli_file outputs for several wavenumber windows
and pyratbay_b200/data/h2o_partition.npz (TIPS H2O partition functions from the TLI header).
"""
import os
import shutil
import subprocess
import sys
import sysconfig

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
WORK = "/tmp/pbref"


def build_scratch():
    pkg = os.path.join(WORK, "pyratbay")
    if not os.path.isdir(pkg):
        os.makedirs(WORK, exist_ok=True)
        shutil.copytree(os.path.join(REF, "pyratbay"), pkg)
    lib = os.path.join(pkg, "lib")
    os.makedirs(lib, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    pyinc = sysconfig.get_paths()["include"]
    if not os.path.exists(os.path.join(pyinc, "Python.h")):
        pyinc = "/usr/include/python3.12"
    npinc = np.get_include()
    for src in sorted(os.listdir(os.path.join(REF, "src_c"))):
        if not src.endswith(".c"):
            continue
        out = os.path.join(lib, src[:-2] + ext)
        if os.path.exists(out):
            continue
        subprocess.check_call([
            "gcc", "-shared", "-fPIC", "-O3", "-ffast-math", "-w", f"-I{pyinc}", f"-I{npinc}",
            f"-I{REF}/src_c/include", os.path.join(REF, "src_c", src), "-o", out, "-lm"])
    sys.path.insert(0, os.path.join(REPO, "oracle", "refstubs"))
    sys.path.insert(0, WORK)


def write_cfg(path, body):
    with open(path, "w") as f:
        f.write("[pyrat]\n" + body)


def main():
    build_scratch()
    import pyratbay as pb
    import pyratbay.opacity as op
    from pyratbay.pyrat import line_by_line as ref_lbl
    import mc3

    run = os.path.join(WORK, "run")
    os.makedirs(os.path.join(run, "outputs"), exist_ok=True)
    os.makedirs(os.path.join(run, "inputs"), exist_ok=True)
    for f in ["Mock_HITRAN_H2O_1.00-1.01um.par", "atmosphere_uniform_test.atm"]:
        shutil.copy(os.path.join(REF, "tests", "inputs", f), os.path.join(run, "inputs", f))
    os.chdir(run)

    # 1. TLI ------------------------------------------------------------------------------
    write_cfg("tli.cfg", """runmode = tli
logfile = outputs/mock.log
dblist = inputs/Mock_HITRAN_H2O_1.00-1.01um.par
dbtype = hitran
pflist = tips
wl_low  = 1.00 um
wl_high = 1.01 um
verb = 1
""")
    pb.run("tli.cfg")
    tli = os.path.join(run, "outputs", "mock.tli")
    shutil.copy(tli, os.path.join(HERE, "mock_hitran_h2o.tli"))

    # TLI reader windows + partition table
    log = mc3.utils.Log(None, verb=0)
    cases = {}
    windows = [(9900.0, 10000.0), (9950.0, 9960.0), (0.0, 1e5), (9999.0, 12000.0),
               (5000.0, 9905.0), (9930.5, 9930.6), (100.0, 200.0)]
    for k, (lo, hi) in enumerate(windows):
        dbs, wn, gf, elow, iso = ref_lbl.read_tli_file(tli, lo, hi, log)
        cases[f"w{k}_range"] = np.array([lo, hi])
        cases[f"w{k}_wn"] = wn
        cases[f"w{k}_gf"] = gf
        cases[f"w{k}_elow"] = elow
        cases[f"w{k}_iso"] = np.asarray(iso)
    np.savez_compressed(os.path.join(HERE, "tli_window_cases.npz"), **cases)
    db = dbs[0]
    os.makedirs(os.path.join(REPO, "pyratbay_b200", "data"), exist_ok=True)
    np.savez_compressed(os.path.join(REPO, "pyratbay_b200", "data", "h2o_partition.npz"),
                        temp=db.temp, z=db.iso_pf, iso_name=np.array(db.iso_name),
                        iso_mass=db.iso_mass, iso_ratio=db.iso_ratio)

    # 2. Opacity table, resample mode -------------------------------------------------------
    base = """runmode = opacity
atmfile = inputs/atmosphere_uniform_test.atm
tlifile = outputs/mock.tli
wl_low   = 1.00 um
wl_high  = 1.01 um
wnosamp = 2160
voigt_extent = 100.0
tmin  =  300
tmax  = 3000
tstep =  300
ncpu = 7
verb = 1
"""
    write_cfg("opacity.cfg", base + "logfile = outputs/table.log\nwnstep = 1.0\n")
    pyrat = pb.run("opacity.cfg")
    ex = pyrat.ex
    atm = pyrat.atm
    np.savez_compressed(
        os.path.join(HERE, "mock_atmosphere.npz"), press=atm.press, temp=atm.temp,
        vmr=atm.vmr, species=np.array(atm.species), mol_mass=atm.mol_mass,
        mol_radius=atm.mol_radius, d=atm.d)
    np.savez_compressed(
        os.path.join(HERE, "mock_opacity_table.npz"), etable=ex.etable, temp=ex.temp,
        press=ex.press, wn=ex.wn, z=ex.z, own0=pyrat.spec.own[0], ownstep=pyrat.spec.ownstep,
        onwave=pyrat.spec.onwave, odivisors=pyrat.spec.odivisors)

    v = pyrat.voigt
    stride = 997
    np.savez_compressed(
        os.path.join(HERE, "mock_voigt.npz"), lorentz=v.lorentz, doppler=v.doppler,
        size=v.size, index=v.index, profile_len=len(v.profile), stride=stride,
        profile_strided=v.profile[::stride],
        profile_first=v.profile[v.index[0, 0]:v.index[0, 0] + 2 * v.size[0, 0] + 1],
        profile_mid=v.profile[v.index[40, 10]:v.index[40, 10] + 2 * v.size[40, 10] + 1],
        profile_last=v.profile[v.index[-1, -1]:v.index[-1, -1] + 2 * v.size[-1, -1] + 1],
        profile_sum=np.sum(v.profile), profile_tail=v.profile[-2000:])

    # 3. Forward model: co-added extinction for all layers, and per-species at one layer ------
    lbl = pyrat.opacity.models[pyrat.opacity.models_type.index('lbl')]
    dens = atm.d[:, lbl.mol_index]
    ec_all = np.copy(lbl.calc_extinction_coefficient(atm.temp, dens))
    ec_layer = np.copy(lbl.calc_extinction_coefficient(atm.temp, dens, layer=31))
    ec_skip = np.copy(lbl.calc_extinction_coefficient(atm.temp, dens, skip_mol=['H2O']))
    np.savez_compressed(os.path.join(HERE, "mock_forward.npz"), ec_all=ec_all,
                        ec_layer31=ec_layer, ec_skip=ec_skip, temp=atm.temp, d=atm.d,
                        lbl_str=np.array(str(lbl).replace(str(lbl.tlifile), "['TLI']")))

    # 4. Line_Sample on the table ---------------------------------------------------------------
    ls = op.Line_Sample(ex.sampled_cs[0])
    temp = np.linspace(450.0, 2900.0, ls.nlayers)
    dens1 = atm.d[:, lbl.mol_index]
    np.savez_compressed(
        os.path.join(HERE, "mock_line_sample.npz"), temperature=temp, density=dens1,
        cs=ls.calc_cross_section(temp), cs_per_mol=ls.calc_cross_section(temp, per_mol=True),
        ec=ls.calc_extinction_coefficient(temp, dens1),
        ec_layer=ls.calc_extinction_coefficient(temp, dens1, layer=20))
    shutil.copy(ex.sampled_cs[0], os.path.join(HERE, "mock_opacity_file.npz"))
    # the same table read twice as two isotopologues with a free and a filler ratio
    iso_a, iso_b = os.path.join(run, "cs_H2O_161.npz"), os.path.join(run, "cs_H2O_181.npz")
    shutil.copy(ex.sampled_cs[0], iso_a)
    shutil.copy(ex.sampled_cs[0], iso_b)
    ls2 = op.Line_Sample([iso_a, iso_b], isotope_ratios="161 main fill_heavy\n181 heavy -2.5")
    dens2 = np.tile(dens1, (1, 2))
    np.savez_compressed(
        os.path.join(HERE, "mock_line_sample_iso.npz"), iso_ratios=ls2.iso_ratios,
        pnames=np.array(ls2.pnames), pars=ls2.pars, species=ls2.species,
        ec=ls2.calc_extinction_coefficient(temp, dens2),
        ec_pars=ls2.calc_extinction_coefficient(temp, dens2, pars=[-3.0]),
        iso_ratios_after=ls2.iso_ratios, cs_per_mol=ls2.calc_cross_section(temp, per_mol=True))

    # 4b. Optical depth of the forward-model extinction (next-tier row) ---------------------------
    from pyratbay.opacity.optic_depth import optical_depth as ref_optical_depth
    radius = 7.0e9 + np.linspace(6.0e8, 0.0, atm.nlayers) ** 1.0     # cm, top to bottom
    od = {"radius": radius, "ec": ec_all}
    for name, kwargs in {
            "emission": dict(rt_path="emission"),
            "emission_max": dict(rt_path="emission", maxdepth=10.0, itop=3, ibottom=45),
            "transit": dict(rt_path="transit"),
            "transit_max": dict(rt_path="transit", maxdepth=10.0, itop=2, ibottom=48)}.items():
        rp, depth, ideep, _, _ = ref_optical_depth(extinction=ec_all * 3e4, radius=radius,
                                                    **kwargs)
        od[name + "_depth"] = depth
        od[name + "_ideep"] = np.asarray(ideep)
    rp, depth, ideep, dclear, iclear = ref_optical_depth(
        "transit", ec_all * 3e4, radius=radius, maxdepth=10.0,
        extinction_cloudy=np.full_like(ec_all, 2e-9))
    od.update(patchy_depth=depth, patchy_ideep=np.asarray(ideep), patchy_depth_clear=dclear,
              patchy_ideep_clear=np.asarray(iclear))
    np.savez_compressed(os.path.join(HERE, "mock_optical_depth.npz"), **od)

    # 5. Opacity table, constant-R (linterp) mode -------------------------------------------------
    write_cfg("opacity_R.cfg", base + "logfile = outputs/table_R.log\nresolution = 15000.0\n"
              "wnstep = 1.0\n")
    pyrat_r = pb.run("opacity_R.cfg")
    np.savez_compressed(
        os.path.join(HERE, "mock_opacity_table_R.npz"), etable=pyrat_r.ex.etable,
        temp=pyrat_r.ex.temp, press=pyrat_r.ex.press, wn=pyrat_r.ex.wn)

    # 6. Voigt grid of the reference's own known-answer test (tests/test_str.py:338-366) --------
    write_cfg("voigt.cfg", """runmode = opacity
logfile = outputs/voigt.log
atmfile = inputs/atmosphere_uniform_test.atm
tlifile = outputs/mock.tli
wl_low   = 1.1 um
wl_high  = 1.7 um
wnstep  = 1.0
wnosamp = 2160
voigt_extent = 100.0
ncpu = 7
verb = 1
""")
    p2 = pb.run("voigt.cfg", run_step='init')
    v2 = p2.voigt
    np.savez_compressed(
        os.path.join(HERE, "voigt_h2o_1.1-1.7um.npz"), lorentz=v2.lorentz, doppler=v2.doppler,
        size=v2.size, index=v2.index, profile_len=len(v2.profile), stride=3989,
        profile_strided=v2.profile[::3989], ownstep=p2.spec.ownstep, onwave=p2.spec.onwave,
        profile_first=v2.profile[v2.index[0, 0]:v2.index[0, 0] + 2 * v2.size[0, 0] + 1],
        profile_last=v2.profile[v2.index[-1, -1]:v2.index[-1, -1] + 2 * v2.size[-1, -1] + 1],
        profile_sum=np.sum(v2.profile), voigt_str=np.array(str(v2)))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
