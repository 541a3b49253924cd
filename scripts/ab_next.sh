#!/bin/bash
# The A/B experiments DESIGN.md section 10 queues for the registration kernel, as one GPU call:
#   bash scripts/ab_next.sh            (on the GPU box; build the variants HERE first: see below)
# Variants are compile-time knobs, built on the CPU box before the call with
#   scripts/build_variant.sh posahead   -DESKF_POS_AHEAD=1
#   scripts/build_variant.sh pospf2     -DESKF_POS_PREFETCH=2
#   scripts/build_variant.sh posahead_pf -DESKF_POS_AHEAD=1 -DESKF_POS_PREFETCH=3
# (eskf_lio_b200/lib/variants/*.so travel with the snapshot).
V=eskf_lio_b200/lib/variants
for v in default posahead pospf2 posahead_pf; do
  if [ $v = default ]; then lib=""; else lib=$PWD/$V/$v.so; [ -f "$lib" ] || continue; fi
  echo "== $v"
  ESKF_GPU_LIB=$lib python scripts/ab_align_opts.py --variants 640:4:2,512:4:2 2>&1 | grep us_per | cut -c1-160
done
# source cloud sorted by voxel key on the host (gathers of neighbouring lanes land in neighbouring voxels)
for s in 0 1; do
  echo "== dense_align --sort $s"
  python scripts/dense_align.py --sort $s 2>&1 | tail -1 | cut -c1-300
done
# where do the tag probes of the compact 0.1 m table hit?  (L2 hit rate / DRAM bytes of one launch)
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum
for c in 0 1; do
  ncu --metrics $M --clock-control none -k regex:align_kernel -s 2 -c 1 --csv --log-file gpurun_out/next_ncu_compact$c.csv \
      python scripts/dense_align.py --reps 1 --warmup 2 --compact $c > /dev/null 2>&1
  grep -h "align_kernel" gpurun_out/next_ncu_compact$c.csv | cut -d, -f13- | tr '\n' ' '; echo " (compact $c)"
done
