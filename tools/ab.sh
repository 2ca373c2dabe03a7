run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-50s %.3e cells/s  %.3f ms  e2e %.3f ms  corr %.3f' % (' '.join(sys.argv[1:]), d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['stage_ms_per_step']['corr']))" "$@"; }
run --opt lanes=2
run --opt lanes=3
run --opt lanes=4
run --opt lanes=2 --opt scratch_mb=80
run --opt lanes=4 --opt scratch_mb=80
run --opt lanes=2 --opt scratch_mb=20
run --opt lanes=3 --opt scratch_mb=60 --opt xchunk_mb=24
run --opt lanes=4 --opt scratch_mb=40 --opt xchunk_mb=110
