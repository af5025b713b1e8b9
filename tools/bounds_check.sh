# GPU parity tests against the bounds-checked build (-DOVO_BOUNDS: every computed offset of the SGBM kernels is checked on the
# device and traps when violated).  Build first, here:  python -c "from openvo_b200 import build; build.build_variant('bounds', ['OVO_BOUNDS'])"
#   gpurun --timeout 900 -- 'bash tools/bounds_check.sh'
export OVO_B200_LIB=openvo_b200/lib/variants/bounds.so
timeout 800 python -m pytest tests -m gpu -x -q > gpurun_out/bounds_tests.log 2>&1
echo "rc=$?"; grep -c OVO_BOUNDS gpurun_out/bounds_tests.log; tail -3 gpurun_out/bounds_tests.log
