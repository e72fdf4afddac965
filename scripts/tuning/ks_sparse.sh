for ks in 1 2 4; do PB200_KSPLIT=$ks python bench.py --steps 5 --warmup 3 --no-cpu-baseline --nlines 100000 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('ks$ks', d['config']['nlines'], d['ms_per_step'], d['detail']['strengths_ms'], d['detail']['accumulate_ms'])"; done
