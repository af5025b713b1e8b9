set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; tail -c 600 gpurun_out/bench_r01_final.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_r01_final_ref.json 2> gpurun_out/bench_r01_final_ref.err; tail -c 400 gpurun_out/bench_r01_final_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 3 --seqs 24 --groups 1 --threads 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sgbm_cost|k_sgbm_vert|k_sgbm_horiz|k_orb_fast_nms|k_knn2_partial_b" -s 5 -c 5 -o gpurun_out/r01b_ncu_full -f python bench.py --steps 2 --warmup 3 --seqs 24 --groups 1 --threads 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/r01b*
