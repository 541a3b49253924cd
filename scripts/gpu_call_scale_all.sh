#!/bin/bash
# bench.py at N = 1, 2, 4, 8 back to back on ONE 8-GPU box (what the driver's scaling run does), collected
# into gpurun_out/r2_scale_builder.json:   gpurun --gpus 8 -- bash scripts/gpu_call_scale_all.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,serial --format=csv,noheader > gpurun_out/r2_scale_gpus.txt
for N in 1 2 4 8; do
  EXTRA="--no-cpu-baseline"
  [ "$N" != 8 ] && EXTRA="$EXTRA --no-batch"
  if [ "$N" = 1 ]; then
    timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 $EXTRA --no-frame > gpurun_out/bench_r2s_n$N.json 2> gpurun_out/bench_r2s_n$N.err
  else
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
      bench.py --gpus $N --steps 20 --warmup 5 $EXTRA > gpurun_out/bench_r2s_n$N.json 2> gpurun_out/bench_r2s_n$N.err
  fi
  echo "bench N=$N rc=$?"
done
python - <<'PY'
import json
out = {}
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f'gpurun_out/bench_r2s_n{n}.json').read().splitlines() if l.startswith('{')][-1])
    except Exception as e:
        out[str(n)] = {"error": str(e)}
        continue
    keep = {k: d.get(k) for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'ms_per_gn_iteration',
                                  'scaling', 'gpu_launches', 'parity', 'strong_scaling_in_this_run', 'nccl_baseline', 'weak', 'batch', 'e2e')}
    keep['roofline'] = {k: d['roofline'].get(k) for k in ('achieved', 'peak', 'frac', 'ms_per_gn_iteration', 'kernel_shape')}
    out[str(n)] = keep
json.dump(out, open('gpurun_out/r2_scale_builder.json', 'w'), indent=1)
v1 = out['1'].get('value')
for n in (1, 2, 4, 8):
    o = out[str(n)]
    print(n, o.get('value'), o.get('ms_per_gn_iteration'), 'x%.2f' % (o.get('value', 0) / v1) if v1 else '', (o.get('weak') or {}).get('ms_per_gn_iteration'),
          (o.get('nccl_baseline') or {}).get('ms_per_gn_iteration'), (o.get('e2e') or {}).get('value'), o['roofline'].get('kernel_shape', '')[:30] if 'roofline' in o else o)
PY
