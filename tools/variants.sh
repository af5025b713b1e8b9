for v in "" vert3_pf4 vert3_pf8 vert3_pf2; do
  if [ -n "$v" ]; then export OVO_B200_LIB=$PWD/openvo_b200/lib/variants/$v.so; else unset OVO_B200_LIB; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --seqs 8 --threads 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('variant', '${v:-base}', 'fps %.0f' % d['value'], 'horiz %.3f vert %.3f cost %.3f fast %.3f' % (k['k_sgbm_horiz_t']['ms_per_launch'], k['k_sgbm_vert_t']['ms_per_launch'], k['k_sgbm_cost_t']['ms_per_launch'], k['k_orb_fast']['ms_per_launch']))"
done
unset OVO_B200_LIB
python -m pytest tests -m gpu -x -q -k "orb or fixtures" 2>&1 | tail -2
