#!/bin/bash
O=gpurun_out/r02v10; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_shipped_path.py -m gpu -q --timeout 600 > $O/pytest_shipped.log 2>&1; echo "rc=$?" >> $O/pytest_shipped.log; tail -4 $O/pytest_shipped.log
