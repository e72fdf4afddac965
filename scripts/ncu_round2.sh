#!/usr/bin/env bash
# Round-2 profiling set on one B200 (B200_PROFILING.md recipe): plain run first, then the launch
# list of the same command, then one --set full capture of the two accumulate kernels and the
# strengths kernel.  Results in gpurun_out/ (summaries are committed under profiles/r02*).
set -u
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-forward-detail"
$CMD > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
$CMD > gpurun_out/r02_plain2.json 2> gpurun_out/r02_plain2.err || exit 1
ncu --set full --clock-control none --import-source on \
    -k regex:"accumulate_chunks_kernel|strengths_kernel|accumulate_dense_kernel" -s 22 -c 3 -f \
    -o gpurun_out/prof_r02_final $CMD > gpurun_out/r02_ncu_full.log 2>&1
tail -n 3 gpurun_out/r02_ncu_full.log
