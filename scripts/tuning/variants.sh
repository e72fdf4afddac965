run() { # label, env...
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bv.json 2> gpurun_out/bv.err
  python -c "
import json; d=json.load(open('gpurun_out/bv.json')); print('$1', d['ms_per_step'], d['detail']['strengths_ms'], d['detail']['accumulate_ms'], d['detail']['checksum'])"
}
run chunk
for k in 2 8; do PB200_KSPLIT=$k run chunk_ks$k; done
for v in $VARIANTS; do PB200_LIB=$PWD/pyratbay_b200/$v.so run $v; done
