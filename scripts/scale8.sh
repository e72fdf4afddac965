# N=8 on one box: forward-model bench (weak scaling) and the 1e8-line table build.
set -u
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 600 gpurun_out/bench_n$N.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/table_build.py --nlines 1e8 --gather > gpurun_out/table_1e8_n$N.json 2> gpurun_out/table_1e8_n$N.err
cat gpurun_out/table_1e8_n$N.json
