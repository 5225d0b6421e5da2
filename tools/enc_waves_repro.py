import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
codec = importlib.import_module("fuse-zstd_b200.codec"); corpus = importlib.import_module("fuse-zstd_b200.corpus")
codec.init([0])
n, size, chunk = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
plain = corpus.json_files(int(sys.argv[4]) if len(sys.argv) > 4 else 4100000, n, size, threads=os.cpu_count())
d_src = torch.from_numpy(plain).cuda()
cap = codec.encode_bound(size, chunk)
d_dst = torch.zeros(n * cap, dtype=torch.uint8, device="cuda")
sp = [d_src.data_ptr() + i * size for i in range(n)]; dp = [d_dst.data_ptr() + i * cap for i in range(n)]
dl, st = codec.encode_batch_ptrs(0, sp, [size] * n, dp, [cap] * n, 3, chunk, codec.SRC_DEVICE | codec.DST_DEVICE)
print("encode ok", st.any(), dl[:3])
