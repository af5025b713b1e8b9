run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 10 --warmup 3 $2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$3 dev %.0f e2e %.0f ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"; }
OVO_SELECT_THREADS=4 run 29541 "" "sel4 24x3"
OVO_SELECT_THREADS=2 run 29542 "" "sel2 24x3"
OVO_SELECT_THREADS=8 run 29543 "--seqs 16 --threads 2" "sel8 16x2"
OVO_SELECT_THREADS=3 run 29544 "--seqs 32 --threads 4" "sel3 32x4"
