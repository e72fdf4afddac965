# Round-end measurement set on one B200 (results merged back through gpurun_out/).
set -u
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 300 gpurun_out/bench_final.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 400 gpurun_out/bench_ref.json; echo
python scripts/table_build.py --nlines 1e8 2>/dev/null | grep "^{" > gpurun_out/table_1e8.json; cut -c1-330 gpurun_out/table_1e8.json
python scripts/multi_species.py > gpurun_out/multi_species.json 2> gpurun_out/multi_species.err; tail -c 500 gpurun_out/multi_species.json
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1j.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"accumulate_chunks|strengths|reduce_partials" -s 3 -c 3 -f -o gpurun_out/prof_r1j $CMD > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu2.log
