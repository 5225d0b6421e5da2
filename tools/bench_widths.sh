#!/bin/bash
# Runs bench.py (device-resident leg only) once per execute-kernel choice: tools/bench_widths.sh t256 t512 1 ...
for w in "$@"; do
  FZG_EXEC_W=$w timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-mount > gpurun_out/b_$w.json 2> gpurun_out/b_$w.err
  python - "$w" <<'PY'
import json, sys
w = sys.argv[1]
try:
    d = json.load(open("gpurun_out/b_%s.json" % w))
    print(w, d["value"], "GB/s", d["ms_per_step"], "ms", {k: round(v, 2) for k, v in d["roofline"]["stage_ms"].items() if v > 0.5})
except Exception as e:
    print(w, "failed", e, open("gpurun_out/b_%s.err" % w).read()[-600:])
PY
done
