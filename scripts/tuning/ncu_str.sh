set -u
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --nlines 10000000"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"strengths" -s 1 -c 1 -f -o gpurun_out/prof_r1i_str $CMD > gpurun_out/ncu1.log 2>&1
tail -n 2 gpurun_out/ncu1.log
