run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$1', d['ms_per_step'], d['detail']['strengths_ms'], d['detail']['accumulate_ms'], d['detail']['checksum'])"; }
run default; run default_1e7 "--nlines 10000000 --steps 3"
for v in $VARIANTS; do export PB200_LIB=$PWD/pyratbay_b200/$v.so; run $v; run ${v}_1e7 "--nlines 10000000 --steps 3"; done
