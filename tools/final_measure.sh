#!/bin/bash
# The records of a round: full GPU test suite, the default bench line (+ reference arm), the ncu launch list of the same
# command and one `--set full` capture of the dominant kernel (k_execute + the k_execute_cta<8> tail) for roofline.traffic.
# usage: tools/final_measure.sh <tag>      (outputs under gpurun_out/<tag>_*)
tag=$1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputests.log 2>&1; tail -3 gpurun_out/${tag}_gputests.log
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_ref.err; cut -c1-300 gpurun_out/${tag}_bench_reference_arm.json
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || { tail -5 gpurun_out/${tag}_bench.err; exit 1; }
cut -c1-700 gpurun_out/${tag}_bench.json
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-mount"
$CMD > gpurun_out/${tag}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > /dev/null 2>&1
$CMD > gpurun_out/${tag}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_execute -s 2 -c 2 -o gpurun_out/${tag}_prof_execute $CMD > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
