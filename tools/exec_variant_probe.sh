#!/bin/bash
# For each variant library (fuse-zstd_b200/variants/libfzgpu_<v>.so; "def" = the shipped one): bench line of k_execute<false> on
# 10 000 files + DRAM bytes / duration of the kernel alone on one wave (4 736 files) from ncu metrics.
# usage: tools/exec_variant_probe.sh <FZG_EXEC_W> v1 v2 ...
w=$1; shift
for v in "$@"; do
  if [ $v = def ]; then unset FZG_LIB; else export FZG_LIB=$PWD/fuse-zstd_b200/variants/libfzgpu_$v.so; fi
  echo "== $v"; tools/bench_widths.sh $w
  CMD="python bench.py --files 4736 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-mount"
  FZG_EXEC_W=$w $CMD > gpurun_out/plain_$v.log 2>&1 && FZG_EXEC_W=$w ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_execute -s 1 -c 1 --csv --log-file gpurun_out/m_$v.csv $CMD > /dev/null 2>&1
  grep -E "k_execute" gpurun_out/m_$v.csv | cut -d, -f 13- | tr -d '"' | awk -F, '{printf "   %s %s %s\n", $2, $4, $3}'
done
