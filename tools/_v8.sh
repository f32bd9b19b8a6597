#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_multi_gpu.py -q --timeout 600 -m gpu -x 2>&1 | tail -30 ) > gpurun_out/r2v8_pytest_multi.log 2>&1; tail -25 gpurun_out/r2v8_pytest_multi.log
