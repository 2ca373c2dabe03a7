run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-50s %.3e cells/s  %.3f ms  e2e %.3f ms  %s' % (' '.join(sys.argv[1:]), d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], {k: round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}))" "$@"; }
for uc in 8 12 14 15 16 18 20 24 30 31 32 40 47 64; do run --opt units_per_chunk=$uc; done
run --opt lanes=3 --opt units_per_chunk=15
run --opt lanes=3 --opt units_per_chunk=10
