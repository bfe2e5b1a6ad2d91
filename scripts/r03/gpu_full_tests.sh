#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 60 | cut -c1-1500 > gpurun_out/${tag}_pytest_gpu.txt; tail -n 40 gpurun_out/${tag}_pytest_gpu.txt
