#!/bin/bash
set -u
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q -rf -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
grep -n "^E  " $O/pytest.log | head -12; tail -5 $O/pytest.log
