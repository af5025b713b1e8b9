# A/B of kernel variants built with openvo_b200.build.build_variant (run under gpurun): bench lines per variant
for cfg in F U; do
for v in base br256_16; do
  if [ $v = base ]; then unset OVO_B200_LIB; else export OVO_B200_LIB=openvo_b200/lib/variants/$v.so; fi
  timeout 400 python bench.py --config $cfg --steps 5 --no-cpu-baseline --no-extras > gpurun_out/r2_ab_${cfg}_$v.json 2> gpurun_out/r2_ab_${cfg}_$v.err
done
done
unset OVO_B200_LIB
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t14.log 2>&1
