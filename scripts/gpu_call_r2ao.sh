#!/bin/bash
# round 2, GPU call AO: bench.py both arms on the final code (traffic read from the committed capture of the tuned kernel)
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
tail -c 400 gpurun_out/bench_r2.json
python bench.py --impl reference > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err
tail -c 300 gpurun_out/bench_ref_r2.json
nvidia-smi --query-gpu=serial --format=csv,noheader > gpurun_out/r2_final_gpu.txt
