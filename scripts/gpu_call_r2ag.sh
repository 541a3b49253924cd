#!/bin/bash
# round 2, GPU call AG: CTA shape of the dense kernel on whatever kind of box answers (the pool has two kinds)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,uuid,serial,pstate,clocks.sm,clocks.mem,clocks.max.sm,clocks.max.mem,power.limit,power.draw,temperature.gpu,ecc.mode.current,memory.total --format=csv | tee gpurun_out/r2ag_gpu_$(date +%s).txt
nvidia-smi -q | grep -i -A3 "Clocks Event\|Max Clocks\|Applications Clocks\|Module Power\|GPU Part\|VBIOS\|Product Brand\|Product Arch" | head -60
timeout 900 python scripts/ab_r2.py --voxels 0.1 --shards 8 --reps 7 --cells "align_block=0;align_block=384;align_block=448;align_block=512;align_block=640;align_block=0;align_block=384" --out gpurun_out/r2ag_ab.json > gpurun_out/r2ag_ab.log 2>&1
echo "ab rc=$?"; cut -c100-330 gpurun_out/r2ag_ab.log | tail -9
