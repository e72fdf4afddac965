#!/usr/bin/env python
"""End-to-end `pbay -c opacity` equivalent at scale on one GPU: TLI file on disk -> Pyrat object
(TLI read, line pre-processing on the device, Voigt table) -> compute_opacity (rows computed
chunk by chunk, copied to pinned host memory and appended to the .npz while the next chunk is on
the GPU) -> .npz on disk (reference layout).

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
    ap.add_argument("--workdir", default="/tmp/pb200_e2e")
    ap.add_argument("--nchunks", type=int, default=5)
    args = ap.parse_args()
    nlines = int(args.nlines)

    import torch
    from pyratbay_b200 import io, tli as ptli, workloads
    from pyratbay_b200 import line_by_line, pyrat as pyrat_mod
    from pyratbay_b200.pyrat import Pyrat

    os.makedirs(args.workdir, exist_ok=True)
    w = workloads.table_workload(nlines)
    tli_path = os.path.join(args.workdir, f"h2o_{nlines}.tli")
    t0 = time.time()
    if not os.path.exists(tli_path):
        wn, elow, gf, iso, counts = w.make_lines()
        ptli.write_tli(tli_path, [w.db], [{"wn": wn, "elow": elow, "gf": gf, "iso_id": iso,
                                           "n_lines_iso": counts}],
                       w.inputs["wnlow"], w.inputs["wnhigh"])
        del wn, elow, gf, iso
    t_tli = time.time() - t0
    out = os.path.join(args.workdir, "table.npz")
    inputs = dict(w.inputs, tlifile=[tli_path], sampled_cs=[out])
    torch.cuda.init()
    torch.zeros(1, device="cuda").sum().item()        # CUDA context up before the clock starts

    # phase timers around the pieces of Pyrat.__init__
    phases = {}

    def timed(name, fn):
        def wrapper(*a, **k):
            t = time.time()
            res = fn(*a, **k)
            phases[name] = phases.get(name, 0.0) + time.time() - t
            return res
        return wrapper
    line_by_line.read_tli_file = timed("tli_read_s", line_by_line.read_tli_file)
    eng_cls = pyrat_mod.Engine
    eng_cls.set_lines = timed("set_lines_s", eng_cls.set_lines)
    eng_cls.build_voigt = timed("voigt_s", eng_cls.build_voigt)

    t0 = time.time()
    pyrat = Pyrat(inputs, atm=w.atm)
    torch.cuda.synchronize()
    t_init = time.time() - t0
    t0 = time.time()
    pyrat.compute_opacity(nchunks=args.nchunks)
    t_table = time.time() - t0
    ex = pyrat.ex
    t0 = time.time()
    back = io.read_opacity(out, extract="opacity")
    t_read = time.time() - t0
    same = bool(np.array_equal(back, ex.etable))
    ok = bool(np.isfinite(ex.etable).all() and (ex.etable >= 0).all())
    print(json.dumps({
        "workload": f"pbay -c opacity equivalent: TLI {nlines:.0e} lines -> {ex.ntemp} T x "
                    f"{ex.nlayers} p x {ex.nwave} wn table -> .npz, 1 GPU",
        "tli_bytes": os.path.getsize(tli_path), "make_tli_s_untimed_input": t_tli,
        "pyrat_init_s": t_init, "init_phases": phases, "compute_opacity_and_write_s": t_table,
        "kernels_ms": {k: float(v) for k, v in ex.timing.items()},
        "total_pipeline_s": t_init + t_table, "npz_bytes": os.path.getsize(out),
        "npz_read_back_s": t_read, "npz_equals_table": same, "finite_nonneg": ok,
        "lines_in_window": pyrat.engine.line_stats()["in_window"]}))


if __name__ == "__main__":
    main()
