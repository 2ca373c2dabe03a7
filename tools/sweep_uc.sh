for uc in 0 24 48 96 192 384; do
  echo "== units_per_chunk=$uc"
  timeout 120 python tools/bench_configs.py "config3 e1b+e1c ref" "l1cd" "config4 native" "l2cm" "b1i" units_per_chunk=$uc 2>&1 | cut -c1-40,95-200
done
