#!/bin/bash
# round 2, GPU call AD: 7-neighbour registration with the neighbourhood's key loads batched: parity + the configs[3] sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -m gpu -x -q -k "direct7 or neighbour or neighbor or sweep or seven or linearize" > gpurun_out/r2ad_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2ad_pytest.log
timeout 1500 python scripts/sweep.py --reps 3 --out gpurun_out/r2_sweep.md > gpurun_out/r2ad_sweep.log 2>&1
echo "sweep rc=$?"; head -22 gpurun_out/r2_sweep.md | cut -c1-200
