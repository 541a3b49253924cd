#!/bin/bash
# bench.py exactly as the driver's scaling run launches it, on the GPUs of this box: bash scripts/gpu_call_scale.sh N [tag]
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
fi
echo "bench N=$N rc=$?"; tail -c 1500 gpurun_out/bench_${TAG}_n$N.err | tail -15
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_${TAG}_n$N.json').read().splitlines() if l.startswith('{')][-1])
    keep = {k: d.get(k) for k in ('metric','value','n_gpus','ms_per_step','ms_per_gn_iteration','gpu_launches','parity','strong_scaling_in_this_run','nccl_baseline','weak','batch')}
    keep['e2e'] = d['e2e']; keep['roofline'] = {k: d['roofline'][k] for k in ('frac','ms_per_gn_iteration','achieved')}
    keep['frame'] = d.get('frame') if not d.get('frame') else {k: d['frame'].get(k) for k in ('value','ms_per_frame_per_rank','timed_frames')}
    print(json.dumps(keep, indent=1)[:5000])
except Exception as e:
    print('parse failed', e)
PY
