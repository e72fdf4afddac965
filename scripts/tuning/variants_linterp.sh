for v in "" variant_o3 variant_o5 variant_ou16 variant_ou4; do
  if [ -n "$v" ]; then export PB200_LIB=$PWD/pyratbay_b200/$v.so; else unset PB200_LIB; fi
  python scripts/table_build.py --nlines 1e7 --resolution 20000 --ntemp 4 --nlayers 13 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('linterp $v', d['accumulate_ms'], d['checksum_rank0'])"
  PB200_ACC_KERNEL=owner python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('owner-transposed $v', d['detail']['accumulate_ms'])"
done
