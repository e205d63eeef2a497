#!/bin/bash
# Full GPU test suite (one pytest process per group) followed by the same-box A/B of the step latency against
# build_tmp/libhsidm_head.so (the previous commit), when that file exists.
bash scripts/gpu_check.sh ${1:-suite} 2>&1 | grep -E "^===|^exit|passed|failed|error" 
if [ -f build_tmp/libhsidm_head.so ]; then
  echo "== head";  HSIDM_AB_LIB=build_tmp/libhsidm_head.so timeout 300 python scripts/step_latency.py 2>&1 | tail -4
fi
echo "== tree";  timeout 300 python scripts/step_latency.py 2>&1 | tail -4
