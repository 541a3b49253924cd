#!/bin/bash
# round 2, final 1-GPU call on the final code: round_end (tests, ncu captures, bench both arms, launch list, 1000 frames) + the configs[3] sweep
bash scripts/round_end.sh r2
timeout 900 python scripts/sweep.py --reps 3 --out gpurun_out/r2_sweep.md > gpurun_out/r2_sweep.log 2>&1; echo "sweep rc=$?"
nvidia-smi --query-gpu=uuid,serial --format=csv,noheader > gpurun_out/r2_final_gpu.txt
