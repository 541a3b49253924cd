#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest_gpu.log 2>&1
echo "pytest(all gpu) rc=$?"; tail -15 gpurun_out/r2j_pytest_gpu.log
