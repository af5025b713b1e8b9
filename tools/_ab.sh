timeout 1200 python -m pytest tests -q -m gpu -k "orb or update" 2>&1 | tail -2
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/sw.json 2> gpurun_out/sw.err
python - <<P
import json
d=json.loads(open('gpurun_out/sw.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), {k:round(x['ms_per_launch'],3) for k,x in d['kernels'].items() if 'orb' in k or 'prep' in k})
P
done
