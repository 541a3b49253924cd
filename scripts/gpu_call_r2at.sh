#!/bin/bash
# round 2, GPU call AT: the 1000-frame sequence on the final code
mkdir -p gpurun_out
timeout 600 python scripts/sequence_full.py --out gpurun_out/sequence_1000_r2.json > gpurun_out/sequence_1000_r2.log 2>&1
echo "rc=$?"; python - <<'PY'
import json
s = json.load(open('gpurun_out/sequence_1000_r2.json'))
print(s['device_ms_per_frame'], s['e2e_wall_ms_per_frame'], s['frames_over_1ms_device'], s['trajectory_match_vs_cpu'])
print([ (f['frame'], round(f['device_ms'],3), f['gn_iterations'], f['eviction_sweep']) for f in s['slowest_frames'][:6]])
PY
