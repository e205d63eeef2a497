#!/bin/bash
# Tile-order A/B of the halo kernel: CTAs per contiguous-range group (1 = each CTA its own range, 148 = grid-wide stride).
for g in 1 2 4 8 16 37 148; do
  echo "== HSIDM_TILE_GROUP=$g"; STEP_LAT_N=176 HSIDM_TILE_GROUP=$g timeout 300 python scripts/step_latency.py 2>&1 | tail -1
done
echo "== group 148 + finalize launches"; STEP_LAT_N=176 HSIDM_TILE_GROUP=148 HSIDM_VARIANT=256 timeout 300 python scripts/step_latency.py 2>&1 | tail -1
echo "== group 4 + finalize launches"; STEP_LAT_N=176 HSIDM_TILE_GROUP=4 HSIDM_VARIANT=256 timeout 300 python scripts/step_latency.py 2>&1 | tail -1
echo "== small batches, group 4"; STEP_LAT_N=1,5,11 HSIDM_TILE_GROUP=4 timeout 300 python scripts/step_latency.py 2>&1 | tail -3
echo "== small batches, group 1"; STEP_LAT_N=1,5,11 HSIDM_TILE_GROUP=1 timeout 300 python scripts/step_latency.py 2>&1 | tail -3
