#!/bin/bash
timeout 300 python -m pytest -q -x -p no:cacheprovider tests/test_e2e_gpu.py -k "validation_driver" 2>&1 | tail -25
