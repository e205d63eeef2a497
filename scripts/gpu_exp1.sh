#!/bin/bash
mkdir -p gpurun_out/exp1
python scripts/layer_prof.py --out gpurun_out/exp1/base.csv --top 12 > gpurun_out/exp1/base.txt 2>&1; head -14 gpurun_out/exp1/base.txt
python scripts/layer_prof.py --variant 16 --out gpurun_out/exp1/v16.csv > gpurun_out/exp1/v16.txt 2>&1; grep -E "total|BN64" gpurun_out/exp1/v16.txt | head -14
HSIDM_NO_RSFUSE_MAX=64 python scripts/layer_prof.py --out gpurun_out/exp1/norsfuse.csv > gpurun_out/exp1/norsfuse.txt 2>&1; grep -E "total|BN64|pertap BN64|cout64" gpurun_out/exp1/norsfuse.txt | head -16
