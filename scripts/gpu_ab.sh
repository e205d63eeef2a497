#!/bin/bash
# Last sanity pass after the schedule-signature change: sampler / end-to-end / drop-in tests and smoke.
timeout 900 python -m pytest -q -x -p no:cacheprovider tests/test_sampler_gpu.py tests/test_e2e_gpu.py tests/test_dropin_gpu.py tests/test_scene_gpu.py -k "not T2000" 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
