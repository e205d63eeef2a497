#!/bin/bash
# ncu --set full of ONE launch: $1 = tag, $2 = kernel regex, $3 = launches to skip. Output: raw + source CSV pages in gpurun_out/$1/.
TAG=$1; KRE=$2; SKIP=${3:-0}
OUT=gpurun_out/$TAG
mkdir -p $OUT /tmp/ncu
CMD="python scripts/step_time.py --precision bf16 --batches 176 --iters 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -o /tmp/ncu/$TAG $CMD > $OUT/ncu.log 2>&1
echo ncu exit $?
ncu -i /tmp/ncu/$TAG.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i /tmp/ncu/$TAG.ncu-rep --page source --csv > $OUT/source.csv 2>/dev/null
ncu -i /tmp/ncu/$TAG.ncu-rep --page details > $OUT/details.txt 2>/dev/null
ls -la $OUT
