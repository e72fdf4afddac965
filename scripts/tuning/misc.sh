run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$1', d['ms_per_step'], d['detail']['strengths_ms'], d['detail']['accumulate_ms'], d['detail']['checksum'])"; }
tab() { python scripts/table_build.py --nlines 1e7 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('table $1', d['strengths_ms'], d['accumulate_ms'])"; }
run default; tab default
for v in $VARIANTS; do export PB200_LIB=$PWD/pyratbay_b200/$v.so; run $v; tab $v; done
