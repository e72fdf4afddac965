# ncu --set full of the accumulate kernel (default lib and one variant), one launch each
set -u
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:accumulate -s 1 -c 1 -f -o gpurun_out/prof_r1f_chunk $CMD > gpurun_out/ncu1.log 2>&1
export PB200_LIB=$PWD/pyratbay_b200/variant_cu16.so
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:accumulate -s 1 -c 1 -f -o gpurun_out/prof_r1f_chunk_u16 $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
