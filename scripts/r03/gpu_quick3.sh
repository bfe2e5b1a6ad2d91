#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_configs.py -m gpu -q -k "mobile" 2>&1 | tail -n 40 | cut -c1-900 > gpurun_out/${tag}_tests_models.txt; tail -n 30 gpurun_out/${tag}_tests_models.txt
