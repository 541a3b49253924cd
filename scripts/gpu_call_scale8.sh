#!/bin/bash
bash scripts/gpu_call_scale.sh 8
echo "== stamps N=8 (rank 0)"
ESKF_ALIGN_STAMPS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
    scripts/dense_sharded.py --reps 2 > gpurun_out/r2_stamps_n8.json 2> gpurun_out/r2_stamps_n8.err
grep "eskf stamps" gpurun_out/r2_stamps_n8.err | tail -12
tail -1 gpurun_out/r2_stamps_n8.json | cut -c1-600
