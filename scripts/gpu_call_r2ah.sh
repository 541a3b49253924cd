#!/bin/bash
# round 2, GPU call AH: why some boxes run the 4-deep 512-thread kernel at ~104 us and the 3-stage 384-thread one at ~92
mkdir -p gpurun_out
python - <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, "SMs", p.multi_processor_count, "mem", p.total_memory)
import ctypes
rt = ctypes.CDLL("libcudart.so.12") if False else None
PY
nvidia-smi --query-gpu=uuid,serial,clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory --format=csv
timeout 900 python scripts/ab_r2.py --voxels 0.1 --reps 7 --cells "align_block=0;align_filter=0;align_block=384;align_block=384,align_depth=4;align_block=256;align_block=768;align_flags=0;l2_persist=0;align_depth=8;align_block=0" --out gpurun_out/r2ah_ab.json > gpurun_out/r2ah_ab.log 2>&1
echo "ab rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2ah_ab.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['cell'], d['us_per_iter'])
PY
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory,clocks_event_reasons.active --format=csv
