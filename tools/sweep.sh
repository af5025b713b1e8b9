for cfg in "16 2" "24 3" "32 2" "32 4" "48 3"; do set -- $cfg
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --seqs $1 --threads $2 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('seqs/threads $1/$2 dev %.0f e2e %.0f ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
