for n in 1000000 10000000 100000000; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline --nlines $n 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print(d['config']['nlines'], d['ms_per_step'], d['detail']['strengths_ms'], d['detail']['accumulate_ms'], d['detail']['checksum'])"; done
