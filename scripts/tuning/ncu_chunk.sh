# ncu --set full: accumulate kernel in table mode (reduced unit count) and at configs[1]
set -u
CMD="python scripts/table_build.py --nlines 3e7 --ntemp 2 --nlayers 12"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"accumulate" -c 1 -f -o gpurun_out/prof_r1h_table $CMD > gpurun_out/ncu1.log 2>&1
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"accumulate" -s 1 -c 1 -f -o gpurun_out/prof_r1h_fwd $CMD > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/plain.log gpurun_out/ncu1.log gpurun_out/ncu2.log
