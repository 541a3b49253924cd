#!/bin/bash
mkdir -p gpurun_out
for n in 2000000 250000 15000; do
  echo "== stamps src $n"
  ESKF_ALIGN_STAMPS=1 timeout 300 python scripts/dense_align.py --src $n --reps 1 --warmup 2 2> gpurun_out/r2o_stamps_$n.txt | cut -c1-150
  tail -3 gpurun_out/r2o_stamps_$n.txt
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest_gpu.log 2>&1
echo "pytest(all gpu) rc=$?"; tail -4 gpurun_out/r2o_pytest_gpu.log
