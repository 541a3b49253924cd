for v in default fold1_b3 fold2_b3 fold4_b2 fold8_b2; do
  for dyn in 1 0; do
    if [ $v = default ]; then lib=""; else lib=$PWD/eskf_lio_b200/lib/variants/$v.so; fi
    echo "== $v dyn=$dyn"
    ESKF_GPU_LIB=$lib ESKF_ALIGN_DYNAMIC=$dyn python scripts/dense_align.py --voxel 0.1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('0.1m', round(d['ms_per_iter']*1e3,1),'us/iter')"
    ESKF_GPU_LIB=$lib ESKF_ALIGN_DYNAMIC=$dyn python scripts/dense_align.py --voxel 0.5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('0.5m', round(d['ms_per_iter']*1e3,1),'us/iter')"
  done
done
