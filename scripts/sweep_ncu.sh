#!/bin/bash
# DRAM / L2 counters of the registration kernel for the cells of the configs[3] sweep
# (run on a GPU box; results land in gpurun_out/sweep_ncu_<voxel>_<mode>.csv).
#   bash scripts/sweep_ncu.sh [src] [map]
SRC=${1:-2000000}; MAP=${2:-10000000}
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
for vox in 0.1 0.25 0.5 1.0; do
  for mode in 1 7; do
    ncu --metrics $M --clock-control none -k regex:align_kernel -s 2 -c 1 --csv \
        --log-file gpurun_out/sweep_ncu_${vox}_${mode}.csv \
        python scripts/dense_align.py --voxel $vox --mode $mode --src $SRC --map $MAP --reps 1 --warmup 2 --hint 11500000 > /dev/null 2>&1
  done
done
