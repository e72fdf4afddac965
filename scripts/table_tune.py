#!/usr/bin/env python
"""Full bench-table batch (1020 units, one GPU) under several settings of the dense/gather split;
prints accumulate / dense kernel times per setting.

    python scripts/table_tune.py "occ=0.25,span=24" "occ=0.1,span=24" ...
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from pyratbay_b200 import constants as pc, workloads
    from pyratbay_b200.engine import Engine
    from pyratbay_b200.spectrum import Spectrum
    from pyratbay_b200.voigt import Voigt
    import torch
    settings = sys.argv[1:] or ["occ=0.25,span=24"]
    nlines = int(float(os.environ.get("NLINES", "1e8")))
    w = workloads.table_workload(nlines)
    resolution = float(os.environ.get("RESOLUTION", "0"))   # constant-R variant of the table
    if resolution:
        spec = Spectrum(wnlow=w.inputs["wnlow"], wnhigh=w.inputs["wnhigh"], wnstep=1.0,
                        resolution=resolution)
    else:
        spec = Spectrum(wnlow=w.inputs["wnlow"], wnhigh=w.inputs["wnhigh"], wnstep=w.wnstep,
                        wnosamp=w.wnosamp)
    lwn, elow, gf, iso, _ = w.make_lines()
    eng = Engine(0)
    eng.set_grid(spec.wn, spec.own, spec.odivisors)
    eng.set_species(w.atm.mol_radius, w.atm.mol_mass, w.iso_atm_index, w.db.iso_mass,
                    w.db.iso_ratio)
    eng.set_lines(lwn, elow, gf, iso.astype(np.int64))
    Voigt(spec, w.atm, w.iso_atm_index, eng, tmin=w.inputs["tmin"], tmax=w.inputs["tmax"])
    n_units = w.ntemp * w.nlayers
    idx = np.arange(n_units)
    itemp, ilayer = idx // w.nlayers, idx % w.nlayers
    temps = w.temps[itemp]
    dens = w.atm.vmr[ilayer] * w.atm.press[ilayer, None] * pc.bar / (pc.k * temps[:, None])
    isoz = workloads.partition(w.db, w.temps)[itemp]
    out = torch.empty((n_units, 1, spec.nwave), dtype=torch.float64, device="cuda:0")
    ref = None
    for setting in settings:
        kv = dict(item.split("=") for item in setting.split(","))
        os.environ["PB200_DENSE"] = kv.get("dense", "1")
        os.environ["PB200_DENSE_MIN_OCC"] = kv.get("occ", "0.25")
        os.environ["PB200_DENSE_MIN_SPAN"] = kv.get("span", "24")
        os.environ["PB200_DYN_FACTOR"] = kv.get("dyn", "0.1")
        if "form" in kv:
            os.environ["PB200_CHUNK_FORM"] = kv["form"]
        else:
            os.environ.pop("PB200_CHUNK_FORM", None)
        for _ in range(2):
            eng.extinction_batch(temps, dens, isoz, w.iso_mol_index, 1, 1e-30, 0,
                                 1 if resolution else 0, out_device_ptr=out.data_ptr())
        t = eng.last_timing()
        rec = {"setting": setting, "accumulate_ms": round(float(t["accumulate_ms"]), 1),
               "dense_ms": round(eng.dense_ms(), 1), "dense_unit_isos": eng.dense_units(),
               "strengths_ms": round(float(t["strengths_ms"]), 1)}
        if ref is None:
            ref = out.clone()
        else:
            rec["maxdiff_vs_first"] = float(((out - ref).abs().amax(dim=-1) / ref.amax(dim=-1)).max())
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
