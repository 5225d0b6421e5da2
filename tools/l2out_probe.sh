run() { # tag lib persist
  if [ "$2" = def ]; then unset FZG_LIB; else export FZG_LIB=$PWD/fuse-zstd_b200/variants/libfzgpu_$2.so; fi
  if [ "$3" = 0 ]; then unset FZG_L2_PERSIST_MB; else export FZG_L2_PERSIST_MB=$3; fi
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-mount 2>gpurun_out/l2out_$1.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', d['value'], d['ms_per_step'], d['roofline']['stage_ms']['execute'])"
  grep "persisting" gpurun_out/l2out_$1.err | head -1
}
run def def 0
run out1_p0 out1 0
run out1_p48 out1 48
run out1_p96 out1 96
run out1h_p96 out1h 96
run def2 def 0
