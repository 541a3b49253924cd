#!/bin/bash
# round 2, GPU call AS: warm-up now exercises the deskewing preprocess: odometry tests + the first frames of a sequence
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_odometry.py tests/test_golden.py tests/test_sensor_log.py -m gpu -x -q > gpurun_out/r2as_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2as_pytest.log
timeout 300 python scripts/frame_probe.py 12 | head -8
