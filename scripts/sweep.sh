#!/usr/bin/env bash
# BASELINE.json configs[4]: sweep the line count (1e5 .. 1e9) and the pressure regime in BOTH
# modes of the hot path -- cross-section table (20 T x 51 p x 1e5 wn, add=0) and forward model
# (81 layers, 0.5-5 um, add=1) -- with bench.py's own options.  N = number of GPUs (default 1;
# N > 1 runs the table mode under torchrun, strong scaling).  One JSON line per point in
# gpurun_out/sweep_n$N.jsonl, a text table at the end.
#   scripts/sweep.sh [N] [max table lines, default 1e9]
set -u
N=${1:-1}
MAXLINES=${2:-1000000000}
out=gpurun_out/sweep_n$N.jsonl
err=gpurun_out/sweep_n$N.err
mkdir -p gpurun_out
: > $out; : > $err
PORT=29650
launch() {   # bench.py arguments...
  if [ "$N" -gt 1 ]; then
    PORT=$((PORT + 1))
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $PORT bench.py --gpus $N --no-cpu-baseline --no-forward-detail "$@" 2>> $err | grep '^{' | tail -1 >> $out
  else
    timeout 1700 python bench.py --no-cpu-baseline --no-forward-detail "$@" 2>> $err | grep '^{' | tail -1 >> $out
  fi
}
for n in 100000 1000000 10000000 100000000 1000000000; do
  [ "$n" -gt "$MAXLINES" ] && continue
  steps=3; [ "$n" -ge 100000000 ] && steps=2
  launch --workload table --nlines $n --steps $steps --warmup 1
  if [ "$n" -le 100000000 ]; then
    launch --workload table --nlines $n --steps $steps --warmup 1 --ptop 1e-6 --pbottom 1e-4
    launch --workload table --nlines $n --steps $steps --warmup 1 --ptop 1e-3 --pbottom 1e-1
    launch --workload table --nlines $n --steps $steps --warmup 1 --ptop 1 --pbottom 100
  fi
  rm -f /tmp/pb200_bench_table_${n}_*.tli
done
if [ "$N" -eq 1 ]; then
  for n in 100000 1000000 10000000 100000000; do
    launch --workload forward --nlines $n --steps 5 --warmup 3
    launch --workload forward --nlines $n --steps 5 --warmup 3 --ptop 1e-6 --pbottom 1e-4
    launch --workload forward --nlines $n --steps 5 --warmup 3 --ptop 1e-3 --pbottom 1e-1
    launch --workload forward --nlines $n --steps 5 --warmup 3 --ptop 1 --pbottom 100
  done
fi
python - "$out" <<'PY'
import json, sys
print("mode     gpus  nlines   p range (bar)        line*layer/s   ms/step    fp64 frac  hbm frac")
for l in open(sys.argv[1]):
    l = l.strip()
    if not l.startswith("{"):
        continue
    d = json.loads(l); c = d["config"]
    mode = "table" if "ntemp" in c else "forward"
    print(f"{mode:8s} {d['n_gpus']:4d}  {c['nlines']:>7.0e}  {str(c['p_bar']):>18s}  {d['value']:.3e}  "
          f"{d['ms_per_step']:10.3f}  {d['roofline_fp64']['frac']:9.3f}  {d['roofline']['frac']:8.4f}")
PY
