#!/bin/bash
# round 2, GPU call AJ: the sortedness pass over the stamps overlapped with the kernels: tests + bench (frame leg e2e)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2aj_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2aj_pytest.log
timeout 900 python bench.py --no-batch --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2aj_bench.json').read().strip().splitlines()[-1])
f = d['frame']
print('frame value', f['value'], 'e2e', f['e2e']['value'], 'p50', f['e2e']['p50_ms'], 'p99', f['e2e']['p99_ms'], 'max', f['e2e']['max_ms'])
print('stage', f['stage_ms'], 'e2e stage', f['e2e']['stage_ms'], 'h2d', f['e2e']['h2d_bytes_per_step'])
print('dense', d['value'], d['roofline']['frac'], d['roofline']['kernel_shape'][:40])
PY
