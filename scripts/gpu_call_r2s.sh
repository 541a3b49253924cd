#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest_gpu.log 2>&1
echo "pytest(all gpu) rc=$?"; tail -4 gpurun_out/r2s_pytest_gpu.log
timeout 900 python scripts/sequence_full.py --frames 150 --no-oracle --out gpurun_out/r2s_seq150.json > gpurun_out/r2s_seq150.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s_seq150.json'))
print(d['device_ms_per_frame'], d['frames_over_1ms_device'], d['slowest_frames'][:3])
PY
