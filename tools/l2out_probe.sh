#!/bin/bash
# Experiment of profiles/r02_notes.md section 10 (negative): output stores of k_execute with an L2 evict_last policy, with and without a
# persisting-L2 set-aside.  Variant libraries first (on the build host):
#   for v in "out1:-DFZ_EXEC_L2OUT=1" "out1h:-DFZ_EXEC_L2OUT=1 -DFZ_EXEC_L2HINT=1"; do n=${v%%:*}; f=${v#*:};
#     nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $f -shared \
#          -o fuse-zstd_b200/variants/libfzgpu_$n.so fuse-zstd_b200/csrc/fz_{decode,encode,api,cache}.cu -lcudart; done
# then on the GPU box: bash tools/l2out_probe.sh   (FZG_LIB selects the library, FZG_L2_PERSIST_MB the set-aside)
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
