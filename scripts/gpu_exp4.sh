#!/bin/bash
timeout 900 python -m pytest -q --tb=short -p no:cacheprovider tests/test_train_gpu.py -s -x 2>&1 | tail -40
