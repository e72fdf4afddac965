run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$1', d['ms_per_step'], d['detail']['strengths_ms'], d['detail']['accumulate_ms'], d['detail']['checksum'])"; }
run default256
export PB200_LIB=$PWD/pyratbay_b200/variant_t512.so
run t512_ks4
PB200_KSPLIT=8 run t512_ks8
PB200_KSPLIT=2 run t512_ks2
python scripts/table_build.py --nlines 1e7 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('table t512', d['accumulate_ms'])"
