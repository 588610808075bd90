#!/bin/bash
# A/B timing of kernel variants: every lib under _variants/ is run through tools/gpu_prof2.py in its own process
for lib in _variants/lib_mprdouble.so _variants/lib_DFB_MPR_FLOAT.so; do
  echo "== $lib"
  FLYBODY_B200_LIB=$PWD/$lib timeout 300 python tools/gpu_prof2.py 2>&1 | grep "ms/step"
done
