#!/usr/bin/env python
"""End-to-end `pbay -c opacity` equivalent at scale on one GPU: TLI file on disk -> Pyrat object
(TLI read, line pre-processing on the device, Voigt table) -> compute_opacity (all (T,p) units
in one batch, table copied to pinned host memory) -> .npz on disk (reference layout).

    python scripts/opacity_e2e.py --nlines 1e8

Prints one JSON line with the wall-clock time of every phase.  The synthetic TLI is written
first (not part of the pipeline being timed)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nlines", type=float, default=1e7)
    ap.add_argument("--nlayers", type=int, default=51)
    ap.add_argument("--nwave", type=int, default=100000)
    ap.add_argument("--wl-low", type=float, default=0.3)
    ap.add_argument("--wl-high", type=float, default=30.0)
    ap.add_argument("--workdir", default="/tmp/pb200_e2e")
    args = ap.parse_args()
    nlines = int(args.nlines)

    import torch
    from pyratbay_b200 import atmosphere as pa, constants as pc, io, tli as ptli, workloads
    from pyratbay_b200.pyrat import Pyrat
    from pyratbay_b200.spectrum import _HCN

    os.makedirs(args.workdir, exist_ok=True)
    wnlow, wnhigh = 1.0 / (args.wl_high * pc.um), 1.0 / (args.wl_low * pc.um)
    wnstep = (wnhigh - wnlow) / (args.nwave - 1)
    wnosamp = int(_HCN[wnstep / _HCN <= 0.0004][0])
    tli_path = os.path.join(args.workdir, f"h2o_{nlines}.tli")
    t0 = time.time()
    ptli.make_synthetic_tli(tli_path, nlines, wnlow, wnhigh, seed=0)
    t_tli = time.time() - t0

    press = pa.pressure(1e-6, 100.0, args.nlayers)
    vmr = np.tile(np.asarray(workloads.UNIFORM_VMR), (args.nlayers, 1))
    atm = pa.Atmosphere(press, np.full(args.nlayers, 1000.0), vmr, workloads.UNIFORM_SPECIES)
    out = os.path.join(args.workdir, "table.npz")
    inputs = dict(tlifile=[tli_path], wnlow=wnlow, wnhigh=wnhigh, wnstep=wnstep, wnosamp=wnosamp,
                  tmin=300.0, tmax=3000.0, tstep=142.1, sampled_cs=[out], verb=0)

    torch.cuda.synchronize()
    t0 = time.time()
    pyrat = Pyrat(inputs, atm=atm)
    torch.cuda.synchronize()
    t_init = time.time() - t0
    t0 = time.time()
    pyrat.compute_opacity()
    t_table = time.time() - t0
    tim = pyrat.last_timing
    ex = pyrat.ex
    t0 = time.time()
    units, species, temp, p, wn, table = None, None, None, None, None, None
    back = io.read_opacity(out)
    t_read = time.time() - t0
    ok = bool(np.isfinite(ex.etable).all() and (ex.etable >= 0).all())
    print(json.dumps({
        "workload": f"pbay -c opacity equivalent: TLI {nlines:.0e} lines -> {ex.ntemp} T x "
                    f"{ex.nlayers} p x {ex.nwave} wn table -> .npz, 1 GPU",
        "tli_bytes": os.path.getsize(tli_path), "make_tli_s_untimed_input": t_tli,
        "pyrat_init_s": t_init, "compute_opacity_s": t_table,
        "kernels_ms": {k: float(v) for k, v in tim.items()},
        "total_pipeline_s": t_init + t_table, "npz_bytes": os.path.getsize(out),
        "npz_read_back_s": t_read, "finite_nonneg": ok,
        "lines_in_window": pyrat.engine.line_stats()["in_window"]}))


if __name__ == "__main__":
    main()
