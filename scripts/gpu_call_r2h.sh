#!/bin/bash
# N = 8: the two exchange protocols, per-rank hand-off stamps
mkdir -p gpurun_out
for ll in 1 0; do
  ESKF_ALIGN_XCHG_LL=$ll timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$ll \
      scripts/dense_sharded.py --reps 5 > gpurun_out/r2h_n8_ll$ll.json 2> gpurun_out/r2h_n8_ll$ll.err
  echo "xchg_ll=$ll rc=$?"; tail -1 gpurun_out/r2h_n8_ll$ll.json | cut -c1-700
done
ESKF_ALIGN_STAMPS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 \
    scripts/dense_sharded.py --reps 1 > gpurun_out/r2h_stamps_n8.json 2> gpurun_out/r2h_stamps_n8.err
grep "eskf stamps" gpurun_out/r2h_stamps_n8.err | grep "it 7" | sort | head -40
