for v in nohoriz2 ""; do
  if [ -n "$v" ]; then export OVO_B200_LIB=$PWD/openvo_b200/lib/variants/$v.so; else unset OVO_B200_LIB; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --seqs 8 --threads 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('variant', '${v:-horiz2}', 'fps %.0f' % d['value'], 'horiz %.3f vert %.3f cost %.3f' % (k['k_sgbm_horiz_t']['ms_per_launch'], k['k_sgbm_vert_t']['ms_per_launch'], k['k_sgbm_cost_t']['ms_per_launch']))"
  python bench.py --steps 12 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   24x3: dev %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"
done
unset OVO_B200_LIB
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
