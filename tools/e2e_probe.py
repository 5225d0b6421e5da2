"""Times fzg_decode_batch with pinned host buffers (the e2e leg of bench.py) with FZG_TRACE per-chunk output."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
codec = importlib.import_module("fuse-zstd_b200.codec")
E = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
S = 1 << 20
w = bench.Workload(0, E, S, 3, os.cpu_count())
codec.init([0])
src_bytes = int(w.comp_off[E - 1] + w.comp_len[E - 1])
h_src = torch.empty(src_bytes + 64, dtype=torch.uint8, pin_memory=True); h_src.numpy()[:src_bytes] = w.packed[:src_bytes]
h_dst = torch.empty(E * S, dtype=torch.uint8, pin_memory=True)
hsp = (h_src.data_ptr() + w.comp_off[:E]).astype(np.uint64)
hdp = (h_dst.data_ptr() + np.arange(E, dtype=np.uint64) * np.uint64(S)).astype(np.uint64)
dc = np.full(E, S, dtype=np.uint64)
for i in range(4):
    t0 = time.perf_counter()
    dl, st = codec.decode_batch_ptrs(0, hsp, w.comp_len[:E], hdp, dc, 0)
    dt = time.perf_counter() - t0
    print("call %d: %.1f ms -> %.1f GB/s e2e, status max %d" % (i, dt * 1e3, E * S / 1e9 / dt, st.max()), file=sys.stderr)
