for v in "" rs64 rs48 rs16; do
  if [ -n "$v" ]; then export OVO_B200_LIB=$PWD/openvo_b200/lib/variants/$v.so; else unset OVO_B200_LIB; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --seqs 8 --threads 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('variant', '${v:-base}', 'fps %.0f' % d['value'], 'horiz %.3f vert %.3f cost %.3f' % (k['k_sgbm_horiz_t']['ms_per_launch'], k['k_sgbm_vert_t']['ms_per_launch'], k['k_sgbm_cost_t']['ms_per_launch']))"
done
