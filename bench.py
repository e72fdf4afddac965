#!/usr/bin/env python
"""Benchmark of the line-by-line hot path (BASELINE.json metric: line x layer
contributions per second).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Workload at N=1 (BASELINE.json configs[1]): synthetic 1e6-line H2O line list, 81-layer
atmosphere (1e-6..100 bar), 0.5-5 um, forward-model extinction (add=1), wnstep=1 cm-1,
wnosamp=2160, Voigt extent 300 HWHM / cutoff 25 cm-1, ethresh 1e-30.  One step = one pass
of the hot path over one atmosphere realisation (81 (T,p) units x all lines): strengths +
accumulate.  At N>1 every rank evaluates its own realisation (weak scaling; the (T,p) units
are independent, there is no data-path collective).

Prints ONE JSON line on rank 0 (contract in the task statement): value = device-resident
throughput (result left in HBM), e2e = the same through the public host API
(Pyrat.calc_lbl_extinction: host arrays in, pinned host array out), roofline for the
accumulate kernel, cpu_baseline = the reference's C path on this box's cores (bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "line_x_layer_contributions_per_sec"
UNIT = "line*layer/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nlines", type=int, default=1_000_000)
    ap.add_argument("--nlayers", type=int, default=81)
    ap.add_argument("--wl-low", type=float, default=0.5, help="um")
    ap.add_argument("--wl-high", type=float, default=5.0, help="um")
    ap.add_argument("--ptop", type=float, default=1e-6, help="bar")
    ap.add_argument("--pbottom", type=float, default=100.0, help="bar")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-layers", type=int, default=0)
    return ap.parse_args()


def workload_config(args):
    return {
        "workload": (f"synthetic {args.nlines:.0e}-line H2O TLI, {args.nlayers}-layer "
                     f"atmosphere {args.ptop:g}..{args.pbottom:g} bar, {args.wl_low:g}-"
                     f"{args.wl_high:g} um forward-model extinction (add=1), "
                     "wnstep=1 cm-1, wnosamp=2160, voigt extent 300 HWHM, cutoff 25 cm-1, "
                     "ethresh 1e-30"),
        "nlines": args.nlines, "nlayers": args.nlayers,
        "wl_um": [args.wl_low, args.wl_high], "p_bar": [args.ptop, args.pbottom],
        "units_per_step_per_gpu": args.nlayers,
        "l2": "no explicit flush: per-step working set (ksum 8 B x groups x layers + Voigt "
              "table 1.2 GB + output) exceeds the 126 MB L2",
    }


# ------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the reference's compiled C path on host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "cpu_baseline.py"),
           "--nlines", str(args.nlines), "--nlayers", str(args.nlayers),
           "--steps", str(args.steps), "--warmup", str(min(args.warmup, 1)),
           "--sample-layers", str(args.cpu_sample_layers)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        print(json.dumps({"impl": "reference", "unavailable":
                          "cpu_baseline.py failed: " + res.stderr.strip()[-300:]}))
        return
    base = json.loads(res.stdout.strip().splitlines()[-1])
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": base["wall_s_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": base["value"], "unit": UNIT, "cores": base["cores"],
                         "kind": base["kind"], "sample": base["sample"],
                         "voigt_grid_s": base["voigt_grid_s"]},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "25", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in out.strip().splitlines():
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(args):
    """DRAM bytes of one accumulate launch from the committed ncu capture; only meaningful for
    the default workload it was taken on."""
    default = (args.nlines == 1_000_000 and args.nlayers == 81 and args.wl_low == 0.5
               and args.wl_high == 5.0 and args.ptop == 1e-6 and args.pbottom == 100.0)
    return 1.491280e9 + 90.537728e6 if default else None


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's banner off stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from pyratbay_b200 import workloads
    from pyratbay_b200.engine import Engine, device_ceilings
    from pyratbay_b200.pyrat import Pyrat
    from pyratbay_b200 import tli as ptli

    # Static set-up (not timed): TLI file -> Pyrat-shaped objects -> engine on this GPU.
    t_setup = time.time()
    w = workloads.forward_model_workload(args.nlines, args.nlayers, args.wl_low, args.wl_high,
                                         ptop=args.ptop, pbottom=args.pbottom)
    tli_path = f"/tmp/pb200_bench_{args.nlines}_{rank}.tli"
    ptli.write_tli(tli_path, [w.db], [{
        "wn": w.wn, "elow": w.elow, "gf": w.gf, "iso_id": w.isoid,
        "n_lines_iso": np.bincount(w.isoid, minlength=w.db.niso)}], w.spec.wnlow, w.spec.wnhigh)
    inputs = dict(tlifile=[tli_path], wl_low=w.spec.wl_low, wl_high=w.spec.wl_high,
                  wnstep=1.0, wnosamp=2160, verb=0)
    pyrat = Pyrat(inputs, atm=w.atm, device=local_rank)
    eng, lbl, atm, spec = pyrat.engine, pyrat.lbl, pyrat.atm, pyrat.spec
    stats = eng.line_stats()
    setup_s = time.time() - t_setup
    nwave, nlayers = spec.nwave, atm.nlayers
    units_per_step = nlayers
    contributions = stats["in_window"] * units_per_step

    stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=torch.device("cuda", local_rank))
    d_out = torch.zeros((nlayers, 1, nwave), dtype=torch.float64, device=f"cuda:{local_rank}")

    def step_inputs(step):
        temps = workloads.layer_temperatures(nlayers, realization=1000 * rank + step)
        atm.calc_profiles(temp=temps)
        return temps, atm.d, workloads.partition(w.db, temps)

    # Per-step host inputs of the device-resident arm are prepared before the timed region
    # (a new atmosphere realisation every step, so nothing can be cached between steps).
    prepared = {s: tuple(np.array(a, copy=True) for a in step_inputs(s))
                for s in range(args.warmup + args.steps)}

    def device_step(step):
        temps, dens, isoz = prepared[step]
        eng.extinction_batch(temps, dens, isoz, lbl.iso_mol_index, lbl.nspec, lbl.ethresh, 1, 0,
                             out_device_ptr=d_out.data_ptr())

    def api_step(step):
        temps = workloads.layer_temperatures(nlayers, realization=1000 * rank + step)
        return pyrat.calc_lbl_extinction(temp=temps)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, first_step):
        """K steps between two events on the engine's stream; max over ranks."""
        for s in range(args.warmup):
            fn(first_step + s)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = eng.launch_count()
        acc_ms, str_ms = [], []
        e0.record(stream)
        for s in range(args.steps):
            fn(first_step + args.warmup + s)
            t = eng.last_timing()
            acc_ms.append(t["accumulate_ms"])
            str_ms.append(t["strengths_ms"])
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), eng.launch_count() - launches0, float(np.mean(acc_ms)), \
            float(np.mean(str_ms))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, launches, acc_ms, str_ms = timed(device_step, 0)
    e2e_ms, _, _, _ = timed(api_step, 100)
    clocks = sampler.stop() if rank == 0 else None

    # Work counters (outside the timed region): surviving groups and gathered samples.
    temps, dens, isoz = step_inputs(0)
    _, cnt = eng.extinction_batch(temps, dens, isoz, lbl.iso_mol_index, lbl.nspec, lbl.ethresh,
                                  1, 0, counters=True, out_device_ptr=d_out.data_ptr())
    neval = int(cnt[:, 2].sum())
    dyn_samples = int(cnt[:, 3].sum())
    gathered = int(cnt[:, 4].sum())
    table_bytes = int(cnt[:, 5].sum())
    checksum = float(d_out.sum().item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / args.steps
    value = contributions * world / (ms_per_step * 1e-3)
    e2e_value = contributions * world / (e2e_ms / args.steps * 1e-3)

    # Roofline of the dominant kernel (accumulate); DESIGN.md section 5 has the derivation.
    # Algorithmic HBM bytes per launch = per (T,p) unit: 20 B per
    # evaluated group (k, head wavenumber, fine index) + 8 B per output sample.  The Voigt
    # samples are gathered out of L2 (roofline_l2); the upper bound of their first-touch HBM
    # bytes is reported separately and is NOT part of `achieved`.
    hbm_peak, peak_src = measured_peaks()
    algo_bytes = 20.0 * neval + 8.0 * nlayers * nwave
    achieved = algo_bytes / (acc_ms * 1e-3) / 1e9
    fp64_tf, l2_gbs = device_ceilings(local_rank)
    roofline = {"bound": "hbm", "kernel": "accumulate_chunks_kernel",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": ncu_traffic(args),
                "traffic_source": "profiles/r01j_final.txt (ncu --set full, "
                                  "dram__bytes_read+write of one launch; default workload only)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes,
                "voigt_table_first_touch_upper_bound_bytes": float(table_bytes),
                "launch_ms": acc_ms, "share_of_step": acc_ms / ms_per_step,
                "note": "HBM traffic equals the algorithmic bytes; the kernel is bound by the L1 "
                        "data pipe and L2->SM bandwidth of the profile gathers (DESIGN.md "
                        "section 5), see roofline_l2 / roofline_fp64 for those ceilings"}
    roofline_l2 = {"bound": "l2", "achieved": 8.0 * gathered / (acc_ms * 1e-3) / 1e9,
                   "peak": l2_gbs, "unit": "GB/s",
                   "frac": 8.0 * gathered / (acc_ms * 1e-3) / 1e9 / l2_gbs,
                   "peak_source": "measured here: 32 MiB L2-resident read loop",
                   "algorithmic_bytes_per_launch": 8.0 * gathered}
    fp64_ach = 2.0 * gathered / (acc_ms * 1e-3) / 1e12
    roofline_fp64 = {"bound": "fp64", "achieved": fp64_ach, "peak": fp64_tf, "unit": "TFLOP/s",
                     "frac": fp64_ach / fp64_tf,
                     "peak_source": "measured here: fp64 FMA microbenchmark",
                     "flops_per_launch": 2.0 * gathered,
                     "reference_equivalent_flops": 2.0 * dyn_samples,
                     "reference_equivalent_tflops": 2.0 * dyn_samples / (acc_ms * 1e-3) / 1e12}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "cpu_baseline.py"),
               "--nlines", str(args.nlines), "--nlayers", str(args.nlayers),
               "--sample-layers", str(args.cpu_sample_layers)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode == 0:
            b = json.loads(res.stdout.strip().splitlines()[-1])
            cpu_baseline = {"value": b["value"], "unit": UNIT, "cores": b["cores"],
                            "kind": b["kind"], "sample": b["sample"],
                            "voigt_grid_s": b["voigt_grid_s"]}
        else:
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                            "sample": "failed: " + res.stderr.strip()[-200:]}

    h2d = 8 * (nlayers * (1 + atm.nmol + lbl.niso)) + 8 * lbl.niso
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 8 * nlayers * nwave, "ms_per_step": e2e_ms / args.steps,
                "api": "Pyrat.calc_lbl_extinction(temp) -> Line_By_Line."
                       "calc_extinction_coefficient (host arrays in, pinned host array out)"},
        "gpu_launches": launches,
        "roofline": roofline, "roofline_l2": roofline_l2, "roofline_fp64": roofline_fp64,
        "cpu_baseline": cpu_baseline,
        "detail": {"lines_in_window": stats["in_window"], "groups": stats["groups"],
                   "nadd": stats["nadd"], "neval_per_step": neval,
                   "dynamic_samples_per_step": dyn_samples, "gathered_samples_per_step": gathered,
                   "strengths_ms": str_ms, "accumulate_ms": acc_ms, "setup_s": setup_s,
                   "voigt_profile_samples": eng.profile_len(), "checksum": checksum},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
