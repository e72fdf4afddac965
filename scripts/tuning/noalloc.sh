run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$1', d['ms_per_step'], d['detail']['strengths_ms'], d['detail']['accumulate_ms'], d['detail']['checksum'])"; }
run default
for v in $VARIANTS; do
PB200_LIB=$PWD/pyratbay_b200/$v.so run $v
PB200_LIB=$PWD/pyratbay_b200/$v.so python scripts/table_build.py --nlines 1e7 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('table $v', d['accumulate_ms'])"
done
