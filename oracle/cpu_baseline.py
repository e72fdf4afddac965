#!/usr/bin/env python
"""TEST/BENCH INFRASTRUCTURE -- times the reference's CPU path on this box's host cores.

Runs the reference's own compiled modules from oracle/_ref (`kind: reference`; falls back
to the C restatement oracle/lbl_oracle.c, `kind: port`, when they were not built) on a
bounded sample of the benchmark workload, with the reference's orchestration: one forked
process per CPU, layers dealt round-robin, results in a shared array
(pyratbay/pyrat/extinction.py:102-119, line_by_line.py:231-246).  Prints one JSON line.

Runs as a separate process from bench.py because fork() and an initialised CUDA context
do not mix.  No CUDA, no pyratbay_b200 engine code on this path (only the host-side grid
builders that define the shared synthetic workload).
"""
import argparse
import ctypes
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def host_cores():
    cores = os.cpu_count() or 2
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    return cores


def table_mode(args):
    """Bounded sample of the cross-section-table workload (BASELINE.json configs[2]): a
    spectral sub-window in the middle of the table's grid (same steps, same line density,
    lines ~ U over the window) x (T,p) units spread evenly over the table's grid, evaluated
    with the reference's orchestration (one forked process per CPU, units dealt round-robin,
    pyratbay/pyrat/extinction.py:100-122) and add=0."""
    import oracle
    from pyratbay_b200 import constants as pc, workloads
    from pyratbay_b200.spectrum import Spectrum
    from pyratbay_b200.voigt import Voigt

    ec_mod, vp_mod = oracle.load_ref()
    kind = "reference"
    if ec_mod is None:
        oracle.build()
        kind = "port"
    cores = host_cores()
    ncpu = args.ncpu if args.ncpu > 0 else max(1, cores - 1)

    count = min(args.window_count, args.nwave)
    first = (args.nwave - count) // 2
    w = workloads.table_workload(args.nlines, args.ntemp, args.nlayers, args.nwave,
                                 window=(first, count))
    spec = Spectrum(wnlow=w.inputs["wnlow"], wnhigh=w.inputs["wnhigh"], wnstep=w.wnstep,
                    wnosamp=w.wnosamp)
    atm, db = w.atm, w.db
    wn, elow, gf, iso, _ = w.make_lines()
    isoid = iso.astype(int)
    v = Voigt(spec, atm, w.iso_atm_index, None, tmin=w.inputs["tmin"], tmax=w.inputs["tmax"])
    profile = np.zeros(v.profile_len, np.double)
    t0 = time.time()
    if kind == "reference":
        vp_mod.grid(profile, v.size, v.index, v.lorentz, v.doppler, spec.ownstep, 0)
    else:
        oracle.grid(profile, v.size, v.index, v.lorentz, v.doppler, spec.ownstep)
    voigt_s = time.time() - t0

    n_units = args.ntemp * args.nlayers
    nsample = args.sample_units if args.sample_units > 0 else min(n_units, 4 * ncpu)
    units = np.unique(np.linspace(0, n_units - 1, nsample).round().astype(int))
    nsample = len(units)
    itemp, ilayer = units // args.nlayers, units % args.nlayers
    temps = w.temps[itemp]
    dens = atm.vmr[ilayer] * atm.press[ilayer, None] * pc.bar / (pc.k * temps[:, None])
    isoz = workloads.partition(db, temps)
    shared = mp.Array(ctypes.c_double, nsample * spec.nwave)
    out = np.ctypeslib.as_array(shared.get_obj()).reshape(nsample, spec.nwave)

    def worker(rank):
        for k in range(rank, nsample, ncpu):
            ext = np.zeros((1, spec.nwave))
            call = (ext, profile, v.size, v.index, v.lorentz, v.doppler, spec.wn, spec.own,
                    spec.odivisors, dens[k], atm.mol_radius, atm.mol_mass, w.iso_atm_index,
                    db.iso_mass, db.iso_ratio, isoz[k], w.iso_mol_index, wn, elow, gf,
                    isoid, v.cutoff, 1e-30, temps[k], 0, 0, 0)
            if kind == "reference":
                ec_mod.extinction(*call)
            else:
                oracle.extinction(*call)
            out[k] = ext[0]

    times = []
    for step in range(args.warmup + args.steps):
        t0 = time.time()
        procs = [mp.get_context('fork').Process(target=worker, args=(r,))
                 for r in range(min(ncpu, nsample))]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        if step >= args.warmup:
            times.append(time.time() - t0)
    wall = float(np.mean(times))
    in_window = int(np.sum((wn >= spec.own[0]) & (wn <= spec.own[-1])))
    print(json.dumps({
        "value": in_window * nsample / wall, "unit": "line*layer/s", "cores": min(ncpu, nsample),
        "host_cores": cores, "kind": kind, "wall_s_per_step": wall, "steps": args.steps,
        "voigt_grid_s": voigt_s, "nlines": in_window, "sample_units": nsample,
        "checksum": float(np.sum(out)),
        "sample": (f"spectral sub-window of the table grid: output samples [{first}, "
                   f"{first + count}) = {spec.wn[0]:.1f}-{spec.wn[-1]:.1f} cm-1 with the table's "
                   f"line density ({in_window} of {args.nlines} lines), {nsample} of {n_units} "
                   f"(T,p) units (evenly spaced), {min(ncpu, nsample)} forked workers; each step "
                   f"= {in_window * nsample:.3g} line x layer in {wall:.2f} s wall; Voigt grid "
                   f"({voigt_s:.1f} s, 1 core) timed separately"),
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nlines", type=int, default=1_000_000)
    ap.add_argument("--nlayers", type=int, default=0, help="0: 81 (forward) / 51 (table)")
    ap.add_argument("--sample-layers", type=int, default=0,
                    help="layers evaluated (0: 2 per worker, at most nlayers)")
    ap.add_argument("--ncpu", type=int, default=0, help="0: host cores - 1 (argum.py:60-66)")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--realization", type=int, default=0)
    ap.add_argument("--mode", default="forward", choices=["forward", "table"])
    ap.add_argument("--ntemp", type=int, default=20)
    ap.add_argument("--nwave", type=int, default=100_000)
    ap.add_argument("--window-count", type=int, default=1000,
                    help="table mode: output samples of the spectral sub-window sampled")
    ap.add_argument("--sample-units", type=int, default=0,
                    help="table mode: (T,p) units evaluated (0: 4 per worker)")
    args = ap.parse_args()
    if args.nlayers <= 0:
        args.nlayers = 51 if args.mode == "table" else 81
    if args.mode == "table":
        return table_mode(args)

    import oracle
    from pyratbay_b200 import workloads
    from pyratbay_b200.voigt import Voigt

    ec_mod, vp_mod = oracle.load_ref()
    kind = "reference"
    if ec_mod is None:
        oracle.build()
        kind = "port"

    cores = host_cores()
    ncpu = args.ncpu if args.ncpu > 0 else max(1, cores - 1)

    w = workloads.forward_model_workload(args.nlines, args.nlayers)
    spec, atm, db = w.spec, w.atm, w.db
    v = Voigt(spec, atm, w.iso_atm_index, None)
    profile = np.zeros(v.profile_len, np.double)
    t0 = time.time()
    if kind == "reference":
        vp_mod.grid(profile, v.size, v.index, v.lorentz, v.doppler, spec.ownstep, 0)
    else:
        oracle.grid(profile, v.size, v.index, v.lorentz, v.doppler, spec.ownstep)
    voigt_s = time.time() - t0

    nsample = args.sample_layers if args.sample_layers > 0 else min(args.nlayers, 2 * ncpu)
    layers = np.unique(np.linspace(0, args.nlayers - 1, nsample).round().astype(int))
    nsample = len(layers)
    shared = mp.Array(ctypes.c_double, nsample * spec.nwave)
    out = np.ctypeslib.as_array(shared.get_obj()).reshape(nsample, spec.nwave)

    def worker(rank, temps, dens, isoz):
        for k in range(rank, nsample, ncpu):
            il = layers[k]
            ext = np.zeros((1, spec.nwave))
            call = (ext, profile, v.size, v.index, v.lorentz, v.doppler, spec.wn, spec.own,
                    spec.odivisors, dens[il], atm.mol_radius, atm.mol_mass, w.iso_atm_index,
                    db.iso_mass, db.iso_ratio, isoz[il], w.iso_mol_index, w.wn, w.elow, w.gf,
                    w.isoid, v.cutoff, 1e-30, temps[il], 0, 1, 0)
            if kind == "reference":
                ec_mod.extinction(*call)
            else:
                oracle.extinction(*call)
            out[k] = ext[0]

    times = []
    for step in range(args.warmup + args.steps):
        temps = workloads.layer_temperatures(args.nlayers, args.realization + step)
        atm.calc_profiles(temp=temps)
        isoz = workloads.partition(db, temps)
        t0 = time.time()
        procs = [mp.get_context('fork').Process(target=worker, args=(r, temps, atm.d, isoz))
                 for r in range(min(ncpu, nsample))]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        if step >= args.warmup:
            times.append(time.time() - t0)
    wall = float(np.mean(times))
    in_window = int(np.sum((w.wn >= spec.own[0]) & (w.wn <= spec.own[-1])))
    print(json.dumps({
        "value": in_window * nsample / wall, "unit": "line*layer/s", "cores": min(ncpu, nsample),
        "host_cores": cores, "kind": kind, "wall_s_per_step": wall, "steps": args.steps,
        "voigt_grid_s": voigt_s, "nlines": in_window, "sample_layers": nsample,
        "checksum": float(np.sum(out)),
        "sample": (f"{nsample} of {args.nlayers} layers (evenly spaced), all {in_window} lines, "
                   f"{min(ncpu, nsample)} forked workers; Voigt grid ({voigt_s:.1f} s, 1 core) "
                   "timed separately"),
    }))


if __name__ == "__main__":
    main()
