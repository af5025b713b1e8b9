for t in 1 2 4 8; do
OVO_SELECT_THREADS=$t python bench.py --steps 12 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('select threads $t dev %.0f e2e %.0f ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
