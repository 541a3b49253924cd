#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest_gpu.log 2>&1
echo "pytest(all gpu) rc=$?"; tail -15 gpurun_out/r2i_pytest_gpu.log
timeout 900 python bench.py --no-frame --no-batch --no-cpu-baseline > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2i.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])
PY
