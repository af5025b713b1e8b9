# A/B of kernel variants built with openvo_b200.build.build_variant (run under gpurun): bench lines per variant
for v in ${VARIANTS:-base}; do
  if [ $v = base ]; then unset OVO_B200_LIB; else export OVO_B200_LIB=openvo_b200/lib/variants/$v.so; fi
  timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-extras > gpurun_out/r2_ab_K_$v.json 2> gpurun_out/r2_ab_K_$v.err
done
unset OVO_B200_LIB
