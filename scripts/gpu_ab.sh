#!/bin/bash
# Same-box A/B/A/B: head = s-outer issue order (no runtime switch), tree = k-outer compiled in.
for i in 1 2; do
echo "== head s-outer";  STEP_LAT_N=176 HSIDM_AB_LIB=build_tmp/libhsidm_head.so timeout 300 python scripts/step_latency.py 2>&1 | tail -1
echo "== tree k-outer";  STEP_LAT_N=176 timeout 300 python scripts/step_latency.py 2>&1 | tail -1
done
echo "== head s-outer";  STEP_LAT_N=1,5,11 HSIDM_AB_LIB=build_tmp/libhsidm_head.so timeout 300 python scripts/step_latency.py 2>&1 | tail -3
echo "== tree k-outer";  STEP_LAT_N=1,5,11 timeout 300 python scripts/step_latency.py 2>&1 | tail -3
