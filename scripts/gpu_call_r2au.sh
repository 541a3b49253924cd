#!/bin/bash
# round 2, GPU call AU: host ESKF reset step with the sparse factor on the left: odometry / golden tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_odometry.py tests/test_golden.py -m gpu -x -q > gpurun_out/r2au_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2au_pytest.log
timeout 200 python scripts/frame_probe.py 60 2>/dev/null | tail -20 | awk '{s+=$5; n++} END {printf "mean dev ms over last %d frames: %.4f\n", n, s/n}'
