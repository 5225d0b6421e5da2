for g in 32 128; do
  export FZG_L2_GRAN=$g
  echo "== L2 fetch granularity $g"
  FZG_EXEC_W=1 python bench.py --files 4736 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-mount 2> gpurun_out/gran_$g.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['stage_ms']['execute'])"
  grep -m1 "L2 fetch" gpurun_out/gran_$g.err
  CMD="python bench.py --files 4736 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-mount"
  FZG_EXEC_W=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_execute -s 1 -c 1 --csv --log-file gpurun_out/m_gran$g.csv $CMD > /dev/null 2>&1
  grep k_execute gpurun_out/m_gran$g.csv | awk -F'","' '{print "   ", $(NF-2), $NF}'
done
