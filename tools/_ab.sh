timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_orb_fast_nms -s 1 -c 1 -o gpurun_out/ncu_fast -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --threads 1 --seqs 8 > gpurun_out/ncu_fast.log 2>&1
ls -la gpurun_out/ncu_fast.ncu-rep
