run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-50s %.3e cells/s  %.3f ms  e2e %.3f ms' % (' '.join(sys.argv[1:]), d['value'], d['ms_per_step'], d['e2e']['ms_per_step']))" "$@"; }
for l in 2 3 4; do for uc in 14 19 20 21 28 42; do run --opt lanes=$l --opt units_per_chunk=$uc; done; done
