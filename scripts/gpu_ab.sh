#!/bin/bash
# Final confirmation on the last commit: the whole GPU suite in one process, as the driver runs it, then a 20-step bench.
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-gpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['traffic'], r['traffic_algorithmic'], r['traffic_layer'][:60])"
