# Evidence capture for profiles/ (run under gpurun on one B200): launch list of a short bench run, then `--set full` of one launch
# of each main kernel.  The .ncu-rep files (tens of MB each with imported source) are exported to CSV on the box and removed, so
# that gpurun_out/ stays under gpurun's 64 MiB return limit.
set -x
CMD="python bench.py --steps 2 --warmup 3 --seqs 24 --groups 1 --threads 1 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/r02_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
$CMD > gpurun_out/r02_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_sgbm_cost|k_sgbm_horiz|k_orb_fast_nms|k_orb_survivors|k_orb_blur|k_knn2_partial" --launch-skip 21 -c 7 -f -o /tmp/r02_full_a $CMD > gpurun_out/r02_ncu_full_a.log 2>&1
$CMD > gpurun_out/r02_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_sgbm_vsum" --launch-skip 80 -c 2 -f -o /tmp/r02_full_b $CMD > gpurun_out/r02_ncu_full_b.log 2>&1
ncu -i /tmp/r02_full_a.ncu-rep --page raw --csv > gpurun_out/r02_full_a_raw.csv
ncu -i /tmp/r02_full_b.ncu-rep --page raw --csv > gpurun_out/r02_full_b_raw.csv
ncu -i /tmp/r02_full_b.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_vsum_source.csv
ls -la gpurun_out /tmp/*.ncu-rep
du -sh gpurun_out
