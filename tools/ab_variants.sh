for v in base h10 h12 c4 c6; do
  if [ $v = base ]; then unset OVO_B200_LIB; else export OVO_B200_LIB=openvo_b200/lib/variants/$v.so; fi
  timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-extras > gpurun_out/r2_ab_$v.json 2> gpurun_out/r2_ab_$v.err
done
unset OVO_B200_LIB
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -k "soak" > gpurun_out/r2_t13.log 2>&1
