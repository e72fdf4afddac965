# ncu --set full of the strengths and accumulate kernels of one bench step
set -u
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"accumulate|strengths" -s 2 -c 2 -f -o gpurun_out/prof_r1g $CMD > gpurun_out/ncu1.log 2>&1
tail -n 3 gpurun_out/ncu1.log
