#!/usr/bin/env python
"""Where does the table build spend its time?  Per pressure (all temperatures of the bench
table at that pressure in one batch): accumulate time with the gather kernels only, with the
dense-convolution path for the main isotope, and the dense kernels' own share.

    python scripts/dense_probe.py [--nlines 1e8] [--layers 0,10,20,30,40,50]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nlines", type=float, default=1e8)
    ap.add_argument("--layers", default="0,8,16,24,32,38,44,50")
    ap.add_argument("--spans", default="24")
    args = ap.parse_args()
    from pyratbay_b200 import constants as pc, workloads
    from pyratbay_b200.engine import Engine
    from pyratbay_b200.spectrum import Spectrum
    from pyratbay_b200.voigt import Voigt
    import torch
    w = workloads.table_workload(int(args.nlines))
    spec = Spectrum(wnlow=w.inputs["wnlow"], wnhigh=w.inputs["wnhigh"], wnstep=w.wnstep,
                    wnosamp=w.wnosamp)
    lwn, elow, gf, iso, _ = w.make_lines()
    eng = Engine(0)
    eng.set_grid(spec.wn, spec.own, spec.odivisors)
    eng.set_species(w.atm.mol_radius, w.atm.mol_mass, w.iso_atm_index, w.db.iso_mass,
                    w.db.iso_ratio)
    eng.set_lines(lwn, elow, gf, iso.astype(np.int64))
    Voigt(spec, w.atm, w.iso_atm_index, eng, tmin=w.inputs["tmin"], tmax=w.inputs["tmax"])
    temps = w.temps
    isoz = workloads.partition(w.db, temps)
    out = torch.empty((len(temps), 1, spec.nwave), dtype=torch.float64, device="cuda:0")
    rows = []
    for il in [int(x) for x in args.layers.split(",")]:
        p = w.atm.press[il]
        dens = w.atm.vmr[il] * p * pc.bar / (pc.k * temps[:, None])
        rec = {"layer": il, "p_bar": float(p)}
        settings = [("gather", "0", "24")] + [(f"dense_s{s}", "1", s) for s in args.spans.split(",")]
        ref = None
        for name, dense, span in settings:
            os.environ["PB200_DENSE"] = dense
            os.environ["PB200_DENSE_MIN_SPAN"] = span
            for _ in range(2):
                res = eng.extinction_batch(temps, dens, isoz, w.iso_mol_index, 1, 1e-30, 0, 0,
                                           out_device_ptr=out.data_ptr(),
                                           counters=(name == "gather"))
                cnt = res[1] if res is not None else None
            t = eng.last_timing()
            rec[name + "_acc_ms"] = round(float(t["accumulate_ms"]), 2)
            if dense == "1":
                rec[name + "_dense_ms"] = round(eng.dense_ms(), 2)
                rec[name + "_units"] = eng.dense_units()
                rec[name + "_maxdiff"] = float((out - ref).abs().max() / ref.max())
            else:
                ref = out.clone()
                rec["gathered_samples"] = int(cnt[:, 4].sum())
        rows.append(rec)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
