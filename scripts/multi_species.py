#!/usr/bin/env python
"""BASELINE.json configs[3]: four single-species cross-section tables (the reference forbids
multi-species tables, pyrat/extinction.py:57-62) on a shared (T,p,wn) grid, kept in HBM as
cs_table[4, ntemp, nlayers, nwave] and consumed by the temperature-interpolation kernel
(interp_ec, src_c/_extcoeff.c:367-418) for one atmosphere.  Prints one JSON line with the
build times and the achieved bandwidth of interp_ec (HBM-bound: two table reads per output
sample and species)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# isotopologue masses / abundances: reference pyratbay/data/isotopes.dat:70-73 (CH4),
# :78-86 (CO), :87-99 (CO2), :146-149 (H2O); leading isotopologues only
SPECIES = {
    "H2O": (["116", "118", "117", "126"], [18.010560, 20.014810, 19.014780, 19.016740],
            [0.997317300, 0.001999827, 0.000371884, 0.000310693]),
    "CH4": (["211", "311", "212"], [16.031300, 17.034655, 17.037475],
            [0.988274000, 0.011103100, 0.000615751]),
    "CO": (["26", "36", "28"], [27.994915, 28.998270, 29.999161],
           [0.986544000, 0.011083600, 0.001978220]),
    "CO2": (["626", "636", "628"], [43.989830, 44.993185, 45.994076],
            [0.984204000, 0.011057400, 0.003947070]),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nlines", type=float, default=1e7)
    ap.add_argument("--ntemp", type=int, default=20)
    ap.add_argument("--nlayers", type=int, default=51)
    ap.add_argument("--nwave", type=int, default=100000)
    args = ap.parse_args()
    nlines = int(args.nlines)

    import torch
    from pyratbay_b200 import atmosphere as pa, constants as pc, tli as ptli, workloads
    from pyratbay_b200.engine import Engine, interp_ec_device
    from pyratbay_b200.spectrum import Spectrum, _HCN
    from pyratbay_b200.voigt import Voigt

    wnlow, wnhigh = 1.0 / (30.0 * pc.um), 1.0 / (0.3 * pc.um)
    wnstep = (wnhigh - wnlow) / (args.nwave - 1)
    spec = Spectrum(wnlow=wnlow, wnhigh=wnhigh, wnstep=wnstep,
                    wnosamp=int(_HCN[wnstep / _HCN <= 0.0004][0]))
    press = pa.pressure(1e-6, 100.0, args.nlayers)
    vmr = np.tile(np.asarray(workloads.UNIFORM_VMR), (args.nlayers, 1))
    atm = pa.Atmosphere(press, np.full(args.nlayers, 1000.0), vmr, workloads.UNIFORM_SPECIES)
    temps = np.linspace(300.0, 3000.0, args.ntemp)
    pf_t, pf_z = ptli.h2o_partition_table()
    dev = torch.device("cuda", 0)
    table = torch.empty((len(SPECIES), args.ntemp, args.nlayers, spec.nwave),
                        dtype=torch.float64, device=dev)
    n_units = args.ntemp * args.nlayers
    itemp, ilayer = np.arange(n_units) // args.nlayers, np.arange(n_units) % args.nlayers
    unit_t = temps[itemp]
    dens = atm.vmr[ilayer] * press[ilayer, None] * pc.bar / (pc.k * unit_t[:, None])
    build = {}
    for s, (name, (iso_names, masses, ratios)) in enumerate(SPECIES.items()):
        niso = len(iso_names)
        db = ptli.Database(f"Synthetic {name}", name, pf_t, iso_names, masses, ratios,
                           pf_z[:niso] * (1.0 + 0.1 * s))
        frac = np.array([0.8, 0.15, 0.05, 0.0][:niso])
        frac = frac / frac.sum()
        wn, elow, gf, iso, _ = ptli.synthetic_lines(nlines, spec.wnlow, spec.wnhigh,
                                                    fractions=tuple(frac), seed=100 + s)
        t0 = time.time()
        eng = Engine(0)
        eng.set_grid(spec.wn, spec.own, spec.odivisors)
        imol = np.full(niso, workloads.UNIFORM_SPECIES.index(name), int)
        eng.set_species(atm.mol_radius, atm.mol_mass, imol, db.iso_mass, db.iso_ratio)
        eng.set_lines(wn, elow, gf, iso.astype(np.int64))
        Voigt(spec, atm, imol, eng, tmin=300.0, tmax=3000.0)
        z = workloads.partition(db, temps)
        torch.cuda.synchronize()
        t1 = time.time()
        eng.extinction_batch(unit_t, dens, z[itemp], np.zeros(niso, int), 1, 1e-30, 0, 0,
                             out_device_ptr=table[s].data_ptr())
        torch.cuda.synchronize()
        build[name] = {"setup_s": t1 - t0, "build_s": time.time() - t1,
                       "groups": eng.line_stats()["groups"]}
        eng.close()
        del eng

    # consume: extinction of one atmosphere from the resident tables
    layer_t = workloads.layer_temperatures(args.nlayers)
    layer_d = pa.ideal_gas_density(vmr, press, layer_t)
    idx = [workloads.UNIFORM_SPECIES.index(n) for n in SPECIES]
    density = np.ascontiguousarray(layer_d[:, idx])
    ext = torch.zeros((args.nlayers, spec.nwave), dtype=torch.float64, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for rep in range(8):
        ext.zero_()
        torch.cuda.synchronize()
        ev0.record()
        interp_ec_device(ext.data_ptr(), table.data_ptr(), temps, layer_t, density, table.shape,
                         0, args.nlayers, stream=torch.cuda.current_stream().cuda_stream)
        ev1.record()
        torch.cuda.synchronize()
        times.append(ev0.elapsed_time(ev1))
    best = min(times[2:])
    nspec = len(SPECIES)
    algo = 8.0 * args.nlayers * spec.nwave * (2 * nspec + 2)   # 2 reads/species + rw of ext
    # check three layers against NumPy
    tab = table[:, :, [0, 25, args.nlayers - 1], :].cpu().numpy()
    got = ext[[0, 25, args.nlayers - 1]].cpu().numpy()
    worst = 0.0
    for k, lay in enumerate([0, 25, args.nlayers - 1]):
        lo = min(int(np.searchsorted(temps, layer_t[lay], side="right")) - 1, args.ntemp - 2)
        w1 = (temps[lo + 1] - layer_t[lay]) / (temps[lo + 1] - temps[lo])
        w2 = (layer_t[lay] - temps[lo]) / (temps[lo + 1] - temps[lo])
        want = sum((tab[j, lo, k] * w1 + tab[j, lo + 1, k] * w2) * density[lay, j]
                   for j in range(nspec))
        worst = max(worst, float(np.max(np.abs(got[k] - want)) / np.max(want)))
    print(json.dumps({
        "workload": f"{nspec} species x {nlines:.0e} lines, {args.ntemp} T x {args.nlayers} p x "
                    f"{spec.nwave} wn, 0.3-30 um", "tables": build,
        "table_bytes_in_hbm": table.numel() * 8,
        "interp_ec_ms": best, "interp_ec_GBps": algo / (best * 1e-3) / 1e9,
        "interp_ec_algorithmic_bytes": algo, "interp_ec_max_rel_err_vs_numpy": worst}))


if __name__ == "__main__":
    main()
