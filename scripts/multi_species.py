#!/usr/bin/env python
"""BASELINE.json configs[3]: four single-species cross-section tables (the reference forbids
multi-species tables, pyrat/extinction.py:57-62) on a shared (T,p,wn) grid, built with
Pyrat.compute_opacity from four synthetic TLI files (units sharded over the ranks when run under
torchrun, rows all-gathered over NCCL), left in HBM, and consumed TOGETHER by one
Line_Sample(tables=[...]) (no .npz round trip) whose temperature interpolation
(src_c/_extcoeff.c:367-418) runs asynchronously into a persistent device buffer.

    python scripts/multi_species.py --nlines 1e8
    torchrun --nproc-per-node 8 scripts/multi_species.py --nlines 1e8

Prints one JSON line (rank 0): build time per table, per-call time and achieved bandwidth of
the interpolation (HBM-bound: two table reads per output sample and species)."""
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
    ap.add_argument("--calls", type=int, default=200)
    args = ap.parse_args()
    nlines = int(args.nlines)

    import torch
    import torch.distributed as dist
    import pyratbay_b200 as pb
    from pyratbay_b200 import atmosphere as pa, tli as ptli, workloads
    from pyratbay_b200.pyrat import Pyrat

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w = workloads.table_workload(nlines)
    pf_t, pf_z = ptli.h2o_partition_table()
    tables, build = [], {}
    for s, (name, (iso_names, masses, ratios)) in enumerate(SPECIES.items()):
        niso = len(iso_names)
        db = ptli.Database(f"Synthetic {name}", name, pf_t, iso_names, masses, ratios,
                           pf_z[:niso] * (1.0 + 0.1 * s))
        frac = np.array([0.8, 0.15, 0.05, 0.0][:niso])
        frac = tuple(frac / frac.sum())
        path = f"/tmp/pb200_multi_{name}_{nlines}.tli"
        if local == 0 and not os.path.exists(path):
            wn, elow, gf, iso, counts = ptli.synthetic_lines(
                nlines, w.inputs["wnlow"], w.inputs["wnlow"] + (w.nwave - 1) * w.wnstep,
                fractions=frac, seed=100 + s)
            ptli.write_tli(path + ".tmp", [db], [{"wn": wn, "elow": elow, "gf": gf, "iso_id": iso,
                                                  "n_lines_iso": counts}],
                           w.inputs["wnlow"], w.inputs["wnhigh"])
            os.replace(path + ".tmp", path)
            del wn, elow, gf, iso
        if world > 1:
            dist.barrier()
        t0 = time.time()
        pyrat = Pyrat(dict(w.inputs, tlifile=[path], sampled_cs=[f"/tmp/pb200_multi_{name}.npz"]),
                      atm=w.atm, device=local)
        torch.cuda.synchronize()
        t1 = time.time()
        pyrat.compute_opacity(write=False, host="none")
        torch.cuda.synchronize()
        t2 = time.time()
        ex = pyrat.ex
        tables.append(dict(name=name, species=name, temp=ex.temp, press=ex.press, wn=ex.wn,
                           opacity=ex.etable_dev.clone()))
        build[name] = {"setup_s": t1 - t0, "build_s": t2 - t1,
                       "groups": pyrat.engine.line_stats()["groups"],
                       "accumulate_ms": ex.timing["accumulate_ms"],
                       "dense_unit_isotopes": ex.timing["dense_units"]}
        ex._assembler = None
        ex.etable_dev = None
        pyrat.engine.close()
        del pyrat, ex
        torch.cuda.empty_cache()

    # consume on every rank: extinction of one atmosphere from the four resident tables
    ls = pb.Line_Sample(tables=tables, device=local)
    del tables
    nlayers, nwave, nspec = ls.nlayers, ls.nwave, ls.nspec
    layer_t = workloads.layer_temperatures(nlayers)
    vmr = np.tile(np.asarray(workloads.UNIFORM_VMR), (nlayers, 1))
    layer_d = pa.ideal_gas_density(vmr, w.atm.press, layer_t)
    idx = [workloads.UNIFORM_SPECIES.index(n) for n in SPECIES]
    density = np.ascontiguousarray(layer_d[:, idx])
    for _ in range(5):
        ext = ls.calc_extinction_coefficient(layer_t, density, device_out=True)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record()
    for _ in range(args.calls):
        ext = ls.calc_extinction_coefficient(layer_t, density, device_out=True)
    ev1.record()
    t_issue = time.time() - t0
    torch.cuda.synchronize()
    per_call_ms = ev0.elapsed_time(ev1) / args.calls
    t0 = time.time()
    host = ls.calc_extinction_coefficient(layer_t, density)
    t_host_call = time.time() - t0
    algo = 8.0 * nlayers * nwave * (2 * nspec + 1)      # 2 reads per species + 1 write
    # check three layers against NumPy
    lays = [0, nlayers // 2, nlayers - 1]
    tab = ls.cs_table_device[:, :, lays, :].cpu().numpy()
    got = ext[lays].cpu().numpy()
    worst = 0.0
    for k, lay in enumerate(lays):
        lo = min(int(np.searchsorted(ls.temp, layer_t[lay], side="right")) - 1, ls.ntemp - 2)
        w1 = (ls.temp[lo + 1] - layer_t[lay]) / (ls.temp[lo + 1] - ls.temp[lo])
        w2 = (layer_t[lay] - ls.temp[lo]) / (ls.temp[lo + 1] - ls.temp[lo])
        want = sum((tab[j, lo, k] * w1 + tab[j, lo + 1, k] * w2) * density[lay, j]
                   for j in range(nspec))
        worst = max(worst, float(np.max(np.abs(got[k] - want)) / np.max(want)))
    assert np.array_equal(host, ext.cpu().numpy())
    if rank == 0:
        print(json.dumps({
            "workload": f"{nspec} species x {nlines:.0e} lines, {ls.ntemp} T x {nlayers} p x "
                        f"{nwave} wn, 0.3-30 um", "n_gpus": world, "tables": build,
            "table_bytes_in_hbm": ls.cs_table_device.numel() * 8,
            "interp_ec_per_call_us_device_out": per_call_ms * 1e3,
            "interp_ec_host_issue_us_per_call": t_issue / args.calls * 1e6,
            "interp_ec_GBps": algo / (per_call_ms * 1e-3) / 1e9,
            "interp_ec_algorithmic_bytes": algo,
            "interp_ec_host_result_call_ms": t_host_call * 1e3,
            "interp_ec_max_rel_err_vs_numpy": worst}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
