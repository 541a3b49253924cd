#!/bin/bash
# round 2, GPU call G: bench.py (N = 1, full) + the two-rank arm on one device
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r2g.err; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_r2g.json').read().strip().splitlines()[-1])
    keep = {k: d[k] for k in ('metric','value','unit','ms_per_step','ms_per_gn_iteration','gpu_launches','clocks','parity')}
    keep['e2e'] = d['e2e']; keep['roofline_frac'] = d['roofline']['frac']; keep['roofline_ms_it'] = d['roofline']['ms_per_gn_iteration']
    keep['cpu'] = d.get('cpu_baseline'); keep['batch'] = d.get('batch')
    f = d.get('frame', {}); keep['frame'] = {k: f.get(k) for k in ('value','e2e','stage_ms','gpu_launches','cpu_baseline','trajectory_match_vs_cpu')}
    keep['single_scan'] = d.get('single_scan')
    print(json.dumps(keep, indent=1)[:6000])
except Exception as e:
    print('parse failed', e)
PY
timeout 900 python -m pytest tests/test_gpu_sharded_p2p.py -m gpu -x -q > gpurun_out/r2g_pytest_sharded.log 2>&1
echo "pytest sharded rc=$?"; tail -15 gpurun_out/r2g_pytest_sharded.log
