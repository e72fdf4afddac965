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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nlines", type=int, default=1_000_000)
    ap.add_argument("--nlayers", type=int, default=81)
    ap.add_argument("--sample-layers", type=int, default=0,
                    help="layers evaluated (0: 2 per worker, at most nlayers)")
    ap.add_argument("--ncpu", type=int, default=0, help="0: host cores - 1 (argum.py:60-66)")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--realization", type=int, default=0)
    args = ap.parse_args()

    import oracle
    from pyratbay_b200 import workloads
    from pyratbay_b200.voigt import Voigt

    ec_mod, vp_mod = oracle.load_ref()
    kind = "reference"
    if ec_mod is None:
        oracle.build()
        kind = "port"

    cores = os.cpu_count() or 2
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
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
