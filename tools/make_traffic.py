"""profiles/traffic.json from an `ncu --set full` report of one bench step: DRAM bytes (read + write) of the execute stage's
kernels (k_execute + the k_execute_cta<8> tail that runs beside its drain), per launch, for bench.py's `roofline.traffic`.
usage: python tools/make_traffic.py gpurun_out/<tag>_prof_execute.ncu-rep <files_per_gpu> [commit] > profiles/traffic.json"""
import csv
import json
import subprocess
import sys

rep, files = sys.argv[1], int(sys.argv[2])
commit = sys.argv[3] if len(sys.argv) > 3 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
parts, total = [], 0.0
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "")
    if "k_execute" not in name:
        continue
    unit_r, unit_w = rows[1][hdr.index("dram__bytes_read.sum")], rows[1][hdr.index("dram__bytes_write.sum")]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    rd = float(d["dram__bytes_read.sum"]) * scale[unit_r]; wr = float(d["dram__bytes_write.sum"]) * scale[unit_w]
    ms = float(d["gpu__time_duration.sum"]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}.get(rows[1][hdr.index("gpu__time_duration.sum")], 1.0)
    parts.append("%s grid %s: read %.3f GB + write %.3f GB, %.2f ms alone" % (name.split("(")[0].replace("void ", ""), d.get("launch__grid_size", "?"), rd / 1e9, wr / 1e9, ms))
    total += rd + wr
print(json.dumps({"kernel": "execute", "files_per_gpu": files, "dram_bytes_per_launch": int(total),
                  "source": "ncu --set full of one step of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-mount` at commit %s (%s): %s" % (commit, rep.split("/")[-1], "; ".join(parts))}, indent=1))
