#!/bin/bash
# round 2, GPU call AA: one-cluster voxelize kernel: stamps after the warp-aggregated histogram + an ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster or downsample or preprocess" > gpurun_out/r2aa_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2aa_pytest.log
ESKF_TRACE=1 timeout 300 python scripts/frame_probe.py 44 2>&1 | grep "voxelize cluster\|preprocess:" | tail -4
timeout 300 python scripts/frame_probe.py 56 > gpurun_out/r2aa_probe.log 2>&1; tail -5 gpurun_out/r2aa_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:voxelize_cluster -s 40 -c 1 -f -o gpurun_out/prof_voxcl_r2 python scripts/frame_probe.py 44 > gpurun_out/r2aa_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_voxcl_r2.ncu-rep
