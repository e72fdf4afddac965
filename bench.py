#!/usr/bin/env python
"""Benchmark of the line-by-line hot path (BASELINE.json metric: line x layer contributions
per second and opacity-table build time at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path
    torchrun --nproc-per-node N bench.py --gpus N ...        # N > 1 (one rank per GPU)

Headline workload (BASELINE.json configs[2], the north-star target): cross-section table of a
synthetic 1e8-line H2O TLI over 20 T x 51 p x 1e5 wavenumbers (0.3-30 um, wnstep 0.33 cm-1,
wnosamp 840, Voigt extent 300 HWHM / cutoff 25 cm-1, ethresh 1e-30, add=0), built through
`Pyrat.compute_opacity` from the TLI file (pyratbay/pyrat/extinction.py:14-126).  One step =
one table build: strengths + accumulate over this rank's share of the 1020 (T,p) units and the
NCCL all-gather that leaves the whole table in the HBM of every rank -- all inside the timed
region.  The SAME table at every N (strong scaling): units are dealt to ranks by estimated
cost, no other data-path collective exists.

Prints ONE JSON line on rank 0 (contract in the task statement): value = line x layer / s with
the result left in HBM, e2e = the same through the public host API (`Pyrat.compute_opacity`:
host arrays in, table in pinned host memory on rank 0 out), roofline of the accumulate kernel,
cpu_baseline = the reference's C path on this box's cores (bounded sample).  `detail` carries
the table build time and, at N = 1, the configs[1] forward-model numbers (`--workload forward`
runs that workload alone, e.g. for the sweeps of scripts/sweep.sh).
"""
import argparse
import gc
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
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="table", choices=["table", "forward"])
    ap.add_argument("--nlines", type=float, default=0, help="0: 1e8 (table) / 1e6 (forward)")
    ap.add_argument("--nlayers", type=int, default=0, help="0: 51 (table) / 81 (forward)")
    ap.add_argument("--ntemp", type=int, default=20)
    ap.add_argument("--nwave", type=int, default=100_000)
    ap.add_argument("--nchunks", type=int, default=4,
                    help="pieces a rank's rows are computed and all-gathered in (N > 1)")
    ap.add_argument("--wl-low", type=float, default=0, help="um; 0: 0.3 (table) / 0.5 (forward)")
    ap.add_argument("--wl-high", type=float, default=0, help="um; 0: 30 (table) / 5 (forward)")
    ap.add_argument("--ptop", type=float, default=1e-6, help="bar")
    ap.add_argument("--pbottom", type=float, default=100.0, help="bar")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-forward-detail", action="store_true")
    ap.add_argument("--cpu-sample-layers", type=int, default=0)
    args = ap.parse_args()
    table = args.workload == "table"
    args.nlines = int(args.nlines) if args.nlines else (100_000_000 if table else 1_000_000)
    args.nlayers = args.nlayers or (51 if table else 81)
    args.wl_low = args.wl_low or (0.3 if table else 0.5)
    args.wl_high = args.wl_high or (30.0 if table else 5.0)
    return args


def table_workload(args):
    from pyratbay_b200 import workloads
    return workloads.table_workload(args.nlines, args.ntemp, args.nlayers, args.nwave,
                                    args.wl_low, args.wl_high, ptop=args.ptop,
                                    pbottom=args.pbottom)


def table_config(args, w):
    return {
        "workload": "BASELINE.json configs[2]: " + w.name + ", voigt extent 300 HWHM, cutoff "
                    "25 cm-1, ethresh 1e-30; built through Pyrat.compute_opacity from the TLI file",
        "nlines": args.nlines, "ntemp": args.ntemp, "nlayers": args.nlayers, "nwave": args.nwave,
        "wl_um": [args.wl_low, args.wl_high], "p_bar": [args.ptop, args.pbottom],
        "units_per_step": args.ntemp * args.nlayers,
        "sharding": "(T,p) units dealt to ranks by estimated cost; NCCL all-gather of the rows "
                    "inside the timed region (chunked, overlapped with compute)",
        "l2": "no explicit flush: the per-step working set (group strengths 8 B x 6.2e7 groups x "
              "20 T = 10 GB, line arrays 2.4 GB, table 0.8 GB) exceeds the 126 MB L2",
    }


def forward_config(args):
    return {
        "workload": (f"BASELINE.json configs[1]: synthetic {args.nlines:.0e}-line H2O TLI, "
                     f"{args.nlayers}-layer atmosphere {args.ptop:g}..{args.pbottom:g} bar, "
                     f"{args.wl_low:g}-{args.wl_high:g} um forward-model extinction (add=1), "
                     "wnstep=1 cm-1, wnosamp=2160, voigt extent 300 HWHM, cutoff 25 cm-1, "
                     "ethresh 1e-30"),
        "nlines": args.nlines, "nlayers": args.nlayers,
        "wl_um": [args.wl_low, args.wl_high], "p_bar": [args.ptop, args.pbottom],
        "units_per_step_per_gpu": args.nlayers,
        "l2": "no explicit flush: per-step working set (ksum 8 B x groups x layers + Voigt "
              "table 1.2 GB + output) exceeds the 126 MB L2",
    }


# ------------------------------------------------------------------------------------------
def cpu_baseline_cmd(args, steps=1, warmup=0):
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "cpu_baseline.py"),
           "--mode", args.workload, "--nlines", str(args.nlines), "--nlayers", str(args.nlayers),
           "--steps", str(steps), "--warmup", str(warmup)]
    if args.workload == "table":
        cmd += ["--ntemp", str(args.ntemp), "--nwave", str(args.nwave)]
    else:
        cmd += ["--sample-layers", str(args.cpu_sample_layers)]
    return cmd


def run_cpu_baseline(args, steps=1, warmup=0):
    res = subprocess.run(cpu_baseline_cmd(args, steps, warmup), capture_output=True, text=True)
    if res.returncode != 0:
        return None, res.stderr.strip()[-300:]
    return json.loads(res.stdout.strip().splitlines()[-1]), None


def run_reference(args):
    """Reference arm: the reference's compiled C path on host cores, bounded sample per step
    (oracle/cpu_baseline.py states the sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warmup = min(args.warmup, 5)     # every step is 2-4 s of wall time on the host cores
    base, err = run_cpu_baseline(args, args.steps, warmup)
    if base is None:
        print(json.dumps({"impl": "reference", "unavailable": "cpu_baseline.py failed: " + err}))
        return
    config = table_config(args, table_workload(args)) if args.workload == "table" \
        else forward_config(args)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup,
        "ms_per_step": base["wall_s_per_step"] * 1e3,
        "ms_per_step_is": "wall time of ONE BOUNDED SAMPLE of the workload (see cpu_baseline."
                          "sample), not of the whole workload; value = sampled line x layer / "
                          "that time",
        "higher_is_better": True,
        "scaling": "strong" if args.workload == "table" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config,
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


def ncu_traffic(workload):
    """DRAM bytes of one accumulate launch of the default workload from the committed ncu
    capture (profiles/TRAFFIC.json: {workload: {"bytes": ..., "source": ...}}), else null."""
    path = os.path.join(ROOT, "profiles", "TRAFFIC.json")
    if os.path.exists(path):
        with open(path) as f:
            ent = json.load(f).get(workload)
        if ent:
            return ent["bytes"], ent["source"]
    return None, None


def rooflines(acc_ms, neval, units, nwave, gathered, dyn_samples, table_bytes, ms_per_step,
              kernel, workload, default_workload, local_rank, launches_per_step=1):
    """roofline objects of the dominant kernel (accumulate); DESIGN.md section 5 has the
    derivation.  Algorithmic HBM bytes per launch = per (T,p) unit 20 B
    per evaluated group (k, head wavenumber, fine index) + 8 B per output sample.  The Voigt
    samples are gathered out of L2 (roofline_l2); the upper bound of their first-touch HBM
    bytes is reported separately and is NOT part of `achieved`.  All quantities are this
    rank's, per launch (a step of a sharded table runs one launch per chunk)."""
    from pyratbay_b200.engine import device_ceilings
    hbm_peak, peak_src = measured_peaks()
    launch_ms = acc_ms / launches_per_step
    algo_bytes = (20.0 * neval + 8.0 * units * nwave) / launches_per_step
    achieved = algo_bytes / (launch_ms * 1e-3) / 1e9
    fp64_tf, l2_gbs = device_ceilings(local_rank)
    traffic, traffic_src = ncu_traffic(workload) if default_workload else (None, None)
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes,
                "voigt_table_first_touch_upper_bound_bytes": float(table_bytes),
                "launch_ms": launch_ms, "launches_per_step": launches_per_step,
                "share_of_step": acc_ms / ms_per_step,
                "note": "HBM traffic is close to the algorithmic bytes; the kernel is bound by "
                        "the delivery of profile samples to the FMA pipe (L1 data pipe / L2->SM), "
                        "see roofline_fp64 / roofline_l2 for those ceilings (DESIGN.md section 5)"}
    l2_ach = 8.0 * gathered / (acc_ms * 1e-3) / 1e9
    roofline_l2 = {"bound": "l2", "achieved": l2_ach, "peak": l2_gbs, "unit": "GB/s",
                   "frac": l2_ach / l2_gbs,
                   "peak_source": "measured here: 32 MiB L2-resident read loop",
                   "algorithmic_bytes_per_step": 8.0 * gathered}
    fp64_ach = 2.0 * gathered / (acc_ms * 1e-3) / 1e12
    roofline_fp64 = {"bound": "fp64", "achieved": fp64_ach, "peak": fp64_tf, "unit": "TFLOP/s",
                     "frac": fp64_ach / fp64_tf,
                     "peak_source": "measured here: fp64 FMA microbenchmark",
                     "flops_per_step": 2.0 * gathered,
                     "note": "flops = 2 x the profile samples that reach an output sample (the "
                             "output-driven count); the reference's formulation evaluates "
                             "reference_equivalent_flops",
                     "reference_equivalent_flops": 2.0 * dyn_samples,
                     "reference_equivalent_tflops": 2.0 * dyn_samples / (acc_ms * 1e-3) / 1e12}
    return roofline, roofline_l2, roofline_fp64


def init_dist():
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
    return world, rank, local_rank


def make_barrier(world):
    import torch
    import torch.distributed as dist

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    return barrier


# ------------------------------------------------------------------------------------------
def forward_model(args, world, rank, local_rank, clocks=True):
    """configs[1]: forward-model extinction of one atmosphere realisation per step (weak
    scaling at N > 1: one realisation per rank).  Returns the result dict on rank 0."""
    import torch
    import torch.distributed as dist
    from pyratbay_b200 import workloads
    from pyratbay_b200.pyrat import Pyrat
    from pyratbay_b200 import tli as ptli

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
    contributions = stats["in_window"] * nlayers

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

    barrier = make_barrier(world)

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
    if rank == 0 and clocks:
        sampler.start()
    total_ms, launches, acc_ms, str_ms = timed(device_step, 0)
    e2e_ms, _, _, _ = timed(api_step, 100)
    clock_info = sampler.stop() if rank == 0 and clocks else None

    # Work counters (outside the timed region): surviving groups and gathered samples.
    temps, dens, isoz = step_inputs(0)
    _, cnt = eng.extinction_batch(temps, dens, isoz, lbl.iso_mol_index, lbl.nspec, lbl.ethresh,
                                  1, 0, counters=True, out_device_ptr=d_out.data_ptr())
    checksum = float(d_out.sum().item())
    if rank != 0:
        return None
    ms_per_step = total_ms / args.steps
    default = (args.nlines == 1_000_000 and args.nlayers == 81 and args.wl_low == 0.5
               and args.wl_high == 5.0 and args.ptop == 1e-6 and args.pbottom == 100.0)
    roofline, roofline_l2, roofline_fp64 = rooflines(
        acc_ms, int(cnt[:, 2].sum()), nlayers, nwave, int(cnt[:, 4].sum()),
        int(cnt[:, 3].sum()), int(cnt[:, 5].sum()), ms_per_step,
        "accumulate_chunks_kernel<4,8>", "forward", default, local_rank)
    h2d = 8 * (nlayers * (1 + atm.nmol + lbl.niso)) + 8 * lbl.niso
    return {
        "value": contributions * world / (ms_per_step * 1e-3), "ms_per_step": ms_per_step,
        "clocks": clock_info, "gpu_launches": launches,
        "e2e": {"value": contributions * world / (e2e_ms / args.steps * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * nlayers * nwave,
                "ms_per_step": e2e_ms / args.steps,
                "api": "Pyrat.calc_lbl_extinction(temp) -> Line_By_Line."
                       "calc_extinction_coefficient (host arrays in, pinned host array out)"},
        "roofline": roofline, "roofline_l2": roofline_l2, "roofline_fp64": roofline_fp64,
        "detail": {"lines_in_window": stats["in_window"], "groups": stats["groups"],
                   "nadd": stats["nadd"], "neval_per_step": int(cnt[:, 2].sum()),
                   "dynamic_samples_per_step": int(cnt[:, 3].sum()),
                   "gathered_samples_per_step": int(cnt[:, 4].sum()),
                   "strengths_ms": str_ms, "accumulate_ms": acc_ms, "setup_s": setup_s,
                   "voigt_profile_samples": eng.profile_len(), "checksum": checksum},
    }


def run_forward(args):
    import torch.distributed as dist
    world, rank, local_rank = init_dist()
    res = forward_model(args, world, rank, local_rank)
    if rank == 0:
        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            b, err = run_cpu_baseline(args)
            cpu_baseline = ({"value": b["value"], "unit": UNIT, "cores": b["cores"],
                             "kind": b["kind"], "sample": b["sample"],
                             "voigt_grid_s": b["voigt_grid_s"]} if b else
                            {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                             "sample": "failed: " + err})
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": forward_config(args), "clocks": res["clocks"],
                "e2e": res["e2e"], "gpu_launches": res["gpu_launches"],
                "roofline": res["roofline"], "roofline_l2": res["roofline_l2"],
                "roofline_fp64": res["roofline_fp64"], "cpu_baseline": cpu_baseline,
                "detail": res["detail"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
def shared_tli(w, args, rank, local_rank, world):
    """One TLI file per box: local rank 0 writes it (atomically), every rank reads it."""
    import torch.distributed as dist
    path = f"/tmp/pb200_bench_table_{args.nlines}_{args.nwave}_{args.wl_low}_{args.wl_high}.tli"
    t0 = time.time()
    if local_rank == 0 and not os.path.exists(path):
        from pyratbay_b200 import tli as ptli
        wn, elow, gf, iso, counts = w.make_lines()
        tmp = path + f".tmp{os.getpid()}"
        ptli.write_tli(tmp, [w.db], [{"wn": wn, "elow": elow, "gf": gf, "iso_id": iso,
                                      "n_lines_iso": counts}],
                       w.inputs["wnlow"], w.inputs["wnhigh"])
        os.replace(tmp, path)
        del wn, elow, gf, iso
    if world > 1:
        dist.barrier()
    return path, time.time() - t0


def run_table(args):
    import torch
    import torch.distributed as dist
    world, rank, local_rank = init_dist()
    from pyratbay_b200 import extinction as ex_mod
    from pyratbay_b200.pyrat import Pyrat

    # Static set-up (not timed): TLI file -> Pyrat-shaped objects -> engine on this GPU.
    w = table_workload(args)
    tli_path, tli_write_s = shared_tli(w, args, rank, local_rank, world)
    t0 = time.time()
    inputs = dict(w.inputs, tlifile=[tli_path],
                  sampled_cs=[f"/tmp/pb200_bench_table_{rank}.npz"])
    pyrat = Pyrat(inputs, atm=w.atm, device=local_rank)
    eng, spec, ex = pyrat.engine, pyrat.spec, pyrat.ex
    stats = eng.line_stats()
    setup_s = time.time() - t0
    n_units = args.ntemp * args.nlayers
    contributions = stats["in_window"] * n_units
    dev = torch.device("cuda", local_rank)
    barrier = make_barrier(world)

    def device_step():
        ex_mod.compute_opacity(pyrat, write=False, host="none", nchunks=args.nchunks)

    def api_step():
        pyrat.compute_opacity(write=False, nchunks=args.nchunks)

    dense_ms = []

    def timed(fn):
        """K table builds between two CUDA events; every build ends with the host blocked on
        the engine's stream and the scatter of the gathered rows queued on torch's current
        stream, where the closing event is recorded.  Max over ranks."""
        for _ in range(args.warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = eng.launch_count()
        acc_ms, str_ms = [], []
        e0.record()
        for _ in range(args.steps):
            fn()
            acc_ms.append(ex.timing["accumulate_ms"])
            str_ms.append(ex.timing["strengths_ms"])
            dense_ms.append(ex.timing["dense_ms"])
        e1.record()
        barrier()
        mine = e0.elapsed_time(e1)
        ms = torch.tensor([mine], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), eng.launch_count() - launches0, float(np.mean(acc_ms)), \
            float(np.mean(str_ms)), mine

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, launches, acc_ms, str_ms, my_ms = timed(device_step)
    e2e_ms, _, _, _, _ = timed(api_step)
    clocks = sampler.stop() if rank == 0 else None

    ex_timing_units = ex.timing["dense_units"]
    table = ex.etable_dev
    finite = bool(torch.isfinite(table).all().item()) and bool((table >= 0).all().item())
    checksum = float(table.sum(dtype=torch.float64).item())
    table_bytes = table.numel() * 8
    # per-rank accumulate / step times (load balance of the strong-scaling split)
    per_rank = torch.tensor([acc_ms, my_ms / args.steps], dtype=torch.float64, device=dev)
    gathered_rank = [torch.zeros_like(per_rank) for _ in range(world)]
    if world > 1:
        dist.all_gather(gathered_rank, per_rank)
    else:
        gathered_rank = [per_rank]
    per_rank_acc = [float(t[0].item()) for t in gathered_rank]
    per_rank_step = [float(t[1].item()) for t in gathered_rank]

    # Work counters of this rank's units (outside the timed region).
    asm = ex._assembler
    mine = asm.mine
    itemp, ilayer = mine // args.nlayers, mine % args.nlayers
    from pyratbay_b200 import constants as pc
    unit_t = ex.temp[itemp]
    dens = pyrat.atm.vmr[ilayer] * pyrat.atm.press[ilayer, None] * pc.bar / (pc.k * unit_t[:, None])
    _, cnt = eng.extinction_batch(unit_t, dens, ex.z[:, itemp].T, pyrat.lbl.iso_mol_index,
                                  pyrat.lbl.nspec, pyrat.lbl.ethresh, 0, 0, counters=True,
                                  out_device_ptr=asm.local.data_ptr())
    # accumulate launches (batches) per device-timed step: one per chunk at N > 1, one at N = 1
    # (the chunked single-rank form only serves the host-output pipelining of the e2e arm)
    launches_per_step = sum(1 for _c, units, _p in asm.chunks() if len(units)) if world > 1 else 1
    nwave = spec.nwave
    voigt_samples = eng.profile_len()
    del table

    forward = None
    if rank == 0 and world == 1 and not args.no_forward_detail:
        # configs[1] (the reference's forward-model case) next to the headline, N = 1 only
        pyrat.ex._assembler = None
        pyrat.ex.etable_dev = None
        del pyrat, eng, asm
        gc.collect()
        torch.cuda.empty_cache()
        fargs = argparse.Namespace(**vars(args))
        fargs.workload, fargs.nlines, fargs.nlayers = "forward", 1_000_000, 81
        fargs.wl_low, fargs.wl_high, fargs.ptop, fargs.pbottom = 0.5, 5.0, 1e-6, 100.0
        fargs.steps, fargs.warmup = max(args.steps, 10), max(args.warmup, 3)
        try:
            forward = forward_model(fargs, 1, 0, local_rank, clocks=False)
            forward["config"] = forward_config(fargs)
        except Exception as exc:  # the headline must not die with the side measurement
            forward = {"error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / args.steps
    value = contributions / (ms_per_step * 1e-3)
    e2e_step = e2e_ms / args.steps
    default = (args.nlines == 100_000_000 and args.ntemp == 20 and args.nlayers == 51
               and args.nwave == 100_000 and args.wl_low == 0.3 and args.wl_high == 30.0
               and world == 1)
    roofline, roofline_l2, roofline_fp64 = rooflines(
        acc_ms, int(cnt[:, 2].sum()), len(mine), nwave, int(cnt[:, 4].sum()),
        int(cnt[:, 3].sum()), int(cnt[:, 5].sum()), ms_per_step,
        "accumulate stage = accumulate_dense_kernel (dense convolution: main isotope + merged "
        "minor isotopes) + accumulate_chunks_kernel<3,16> (gather: the remaining groups)",
        "table", default, local_rank, launches_per_step)
    dense_share = float(np.mean(dense_ms[:args.steps])) / acc_ms if acc_ms > 0 else 0.0
    roofline["dense_kernel_share_of_stage"] = dense_share
    roofline["note"] = (
        "algorithmic HBM bytes of the whole accumulate stage over its device time.  The stage is "
        "compute/delivery bound, not HBM bound (the algorithm needs 1.3 TB per table): the dense "
        "kernel runs on the fp64 FMA pipe (62-66 % active, profiles/r02f_dense_final.txt), the gather "
        "kernel on L2->SM bandwidth; roofline_fp64 is the meaningful ceiling (DESIGN.md section 5)")

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        b, err = run_cpu_baseline(args)
        cpu_baseline = ({"value": b["value"], "unit": UNIT, "cores": b["cores"],
                         "kind": b["kind"], "sample": b["sample"],
                         "voigt_grid_s": b["voigt_grid_s"]} if b else
                        {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                         "sample": "failed: " + err})

    h2d = 8 * n_units * (1 + pyrat_nmol(w) + w.db.niso)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": table_config(args, w),
        "clocks": clocks,
        "e2e": {"value": contributions / (e2e_step * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": table_bytes,
                "ms_per_step": e2e_step,
                "api": "Pyrat.compute_opacity(write=False): host (T,p) arrays in, the table "
                       "[ntemp, nlayers, nwave] in pinned host memory on rank 0 out (the .npz "
                       "write is file I/O, reported by scripts/opacity_e2e.py)"},
        "gpu_launches": launches,
        "roofline": roofline, "roofline_l2": roofline_l2, "roofline_fp64": roofline_fp64,
        "cpu_baseline": cpu_baseline,
        "detail": {"table_build_s": ms_per_step * 1e-3, "table_build_e2e_s": e2e_step * 1e-3,
                   "lines_in_window": stats["in_window"], "groups": stats["groups"],
                   "nadd": stats["nadd"], "units": n_units, "units_rank0": int(len(mine)),
                   "neval_rank0": int(cnt[:, 2].sum()),
                   "dynamic_samples_rank0": int(cnt[:, 3].sum()),
                   "gathered_samples_rank0": int(cnt[:, 4].sum()),
                   "strengths_ms_rank0": str_ms, "accumulate_ms_rank0": acc_ms,
                   "dense_kernels_ms_rank0": float(np.mean(dense_ms[:args.steps])),
                   "dense_unit_isotopes_rank0": int(ex_timing_units),
                   "accumulate_ms_per_rank": per_rank_acc, "step_ms_per_rank": per_rank_step,
                   "allgather_chunks": launches_per_step if world > 1 else 0,
                   "setup_s": setup_s, "tli_write_s": tli_write_s,
                   "voigt_profile_samples": voigt_samples, "finite_nonneg": finite,
                   "checksum": checksum, "forward_model_configs1": forward},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def pyrat_nmol(w):
    return len(w.atm.species)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "forward":
        run_forward(args)
    else:
        run_table(args)


if __name__ == "__main__":
    main()
