#!/usr/bin/env bash
# BASELINE.json configs[4]: sweep line count and pressure regime on one GPU (bench.py options).
# Writes one JSON line per point to gpurun_out/sweep.jsonl.
set -u
out=gpurun_out/sweep.jsonl
: > $out
run() {
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>> gpurun_out/sweep.err | tail -1 >> $out
}
for n in 100000 1000000 10000000; do
  run --nlines $n
  run --nlines $n --ptop 1e-6 --pbottom 1e-4
  run --nlines $n --ptop 1e-3 --pbottom 1e-1
  run --nlines $n --ptop 1 --pbottom 100
done
run --nlines 100000000 --steps 3
python - <<'PY'
import json
print("nlines      p range (bar)      line*layer/s   ms/step  strengths_ms accumulate_ms gathered/group")
for l in open("gpurun_out/sweep.jsonl"):
    l = l.strip()
    if not l.startswith("{"): continue
    d = json.loads(l); c = d["config"]; dd = d["detail"]
    w = c["workload"]
    print(f"{c['nlines']:>10.0e}  {str(c['p_bar']):>18s}  {d['value']:.3e}  {d['ms_per_step']:9.3f}  "
          f"{dd['strengths_ms']:9.3f}  {dd['accumulate_ms']:9.3f}  "
          f"{dd['gathered_samples_per_step']/max(dd['neval_per_step'],1):7.1f}")
PY
