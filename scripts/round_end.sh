#!/bin/bash
# What a round-end GPU call runs (on the GPU box, from the repo root):
#   GPU parity tests -> ncu --set full of the roofline kernel on the bench's dense workload ->
#   its summary into profiles/ (bench.py reads `traffic` from it) -> bench.py -> ncu launch list
#   of a short bench run.  Everything lands in gpurun_out/; scripts/summarize_profiles.py turns it
#   into the committed files under profiles/.
#     bash scripts/round_end.sh [tag]        (default tag: r2)
TAG=${1:-r2}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log; fi
[ -z "$SKIP_TESTS" ] && tail -2 gpurun_out/pytest_gpu_$TAG.log
# (the first registration of the process autotunes: 4 shapes x 4 launches, then 2 warm-ups, then the launch captured)
ncu --set full --clock-control none --import-source on -k regex:align_kernel -s 18 -c 1 -f \
    -o gpurun_out/prof_align_$TAG python scripts/dense_align.py --reps 1 --warmup 2 > gpurun_out/prof_align_$TAG.log 2>&1
[ -n "$SKIP_FRAME_NCU" ] || ncu --set full --clock-control none --import-source on -k regex:voxelize_cluster -s 40 -c 1 -f \
    -o gpurun_out/prof_voxcl_$TAG python scripts/frame_probe.py 44 > gpurun_out/prof_voxcl_$TAG.log 2>&1
python scripts/summarize_profiles.py $TAG > /dev/null 2>&1
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 1200 gpurun_out/bench_$TAG.json
python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -c 600 gpurun_out/bench_ref_$TAG.json
[ -n "$SKIP_LAUNCHES" ] || ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-batch > gpurun_out/bench_under_ncu_$TAG.log 2>&1
[ -n "$SKIP_LAUNCHES" ] || grep -c align_kernel gpurun_out/launches_$TAG.csv
[ -n "$SKIP_SEQ" ] || python scripts/sequence_full.py --out gpurun_out/sequence_1000_$TAG.json > gpurun_out/sequence_1000_$TAG.log 2>&1
[ -n "$SKIP_SEQ" ] || tail -c 900 gpurun_out/sequence_1000_$TAG.log
