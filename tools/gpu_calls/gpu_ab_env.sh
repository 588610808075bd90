#!/bin/bash
# A/B of runtime switches: stage times at 4096 envs with and without each
for v in "" "FB_BLOB=1" "FB_NO_GRAPH=1"; do
  echo "== ${v:-default}"
  env $v timeout 300 python tools/gpu_prof2.py 2>&1 | grep "ms/step"
done
