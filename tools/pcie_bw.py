"""Host link ceiling of the box: H2D / D2H / both directions at once with pinned buffers, on 1..N GPUs CONCURRENTLY
(one process per GPU, a barrier before every timed region), with and without binding each process to its GPU's NUMA node.
What `e2e` of bench.py can reach at most: e2e moves (compressed in) H2D and (plain out) D2H per step.

    python tools/pcie_bw.py [--concurrent N] [--mib 1024] [--reps 5] [--bind 0|1|both] > profiles/r02_pcie_concurrent.json
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time


def gpu_cpus(index):
    """CPUs of the NUMA node GPU `index` hangs off (NVML), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        return cpus or None
    except Exception:
        return None


def worker(rank, n_gpus, mib, reps, bind, barrier, out):
    if bind:
        cpus = gpu_cpus(rank)
        if cpus:
            os.sched_setaffinity(0, cpus)                  # first touch: the pinned buffers below land on this node
    import torch
    torch.cuda.set_device(rank)
    n = mib << 20
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out.fill_(2)
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.zeros(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        h2d(); d2h()

    res = {}
    for name, f, nbytes in (("h2d", h2d, n), ("d2h", d2h, n), ("both", both, 2 * n)):
        for _ in range(2):
            f()
        torch.cuda.synchronize()
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(reps):
            f()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        barrier.wait()
        res[name] = nbytes / 1e9 / dt
    out.put((rank, res, sorted(os.sched_getaffinity(0))[:4]))


def run(n_gpus, mib, reps, bind):
    ctx = mp.get_context("spawn")
    barrier, out = ctx.Barrier(n_gpus), ctx.Queue()
    ps = [ctx.Process(target=worker, args=(r, n_gpus, mib, reps, bind, barrier, out)) for r in range(n_gpus)]
    for p in ps:
        p.start()
    got = [out.get(timeout=600) for _ in ps]
    for p in ps:
        p.join()
    got.sort()
    agg = {k: round(sum(r[1][k] for r in got), 2) for k in ("h2d", "d2h", "both")}
    return {"gpus": n_gpus, "numa_bound": bool(bind), "aggregate_GBps": agg,
            "per_gpu_GBps": [{k: round(v, 2) for k, v in r[1].items()} for r in got], "first_cpus": [r[2] for r in got]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--concurrent", type=int, default=0, help="GPUs used at once (0: 1, 2, 4, ... up to all visible)")
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--bind", default="both", choices=["0", "1", "both"])
    a = ap.parse_args()
    import torch
    vis = torch.cuda.device_count()
    counts = [a.concurrent] if a.concurrent else [c for c in (1, 2, 4, 8) if c <= vis]
    rows = []
    for c in counts:
        for b in ((0, 1) if a.bind == "both" else (int(a.bind),)):
            r = run(c, a.mib, a.reps, b)
            print(json.dumps(r), file=sys.stderr, flush=True)
            rows.append(r)
    print(json.dumps({"tool": "tools/pcie_bw.py", "mib_per_copy": a.mib, "host_cpus": os.cpu_count(), "results": rows}, indent=1))


if __name__ == "__main__":
    main()
