set -u
out=gpurun_out/sweep_multi.jsonl; : > $out
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $N --no-cpu-baseline --no-forward-detail --workload table "$@" 2>> gpurun_out/sweep_multi.err | grep '^{' | tail -1 >> $out; }
run 8 --nlines 10000000 --steps 3 --warmup 1
run 8 --nlines 10000000 --steps 3 --warmup 1 --ptop 1 --pbottom 100
run 8 --nlines 1000000 --steps 3 --warmup 1
run 4 --nlines 10000000 --steps 3 --warmup 1
run 4 --nlines 10000000 --steps 3 --warmup 1 --ptop 1 --pbottom 100
run 2 --nlines 10000000 --steps 3 --warmup 1
python - <<'PY'
import json
for l in open("gpurun_out/sweep_multi.jsonl"):
    if not l.startswith("{"): continue
    d=json.loads(l); c=d["config"]
    print(d["n_gpus"], c["nlines"], c["p_bar"], round(d["ms_per_step"],2), f'{d["value"]:.3e}')
PY
