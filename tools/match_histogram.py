"""Match statistics of the config-2 corpus (1 MiB synthetic JSON files, libzstd level 3, reference-writer framing):
offset / match-length / literal-run distributions and how much of the output a window of the most recent W bytes covers.
TEST / ANALYSIS TOOL (uses the oracle's trace): python tools/match_histogram.py [n_files] > profiles/r02_match_histogram.json
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle

corpus = importlib.import_module("fuse-zstd_b200.corpus")


def main():
    n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    size = 1 << 20
    orc, ref = pyoracle.Oracle(), pyoracle.Ref()
    offs, mls, lls, pos = [], [], [], []
    nseq_blocks = []
    comp = 0
    for i in range(n_files):
        plain = corpus.json_file(i, size).tobytes()
        z = ref.writer_encode(plain, 3)
        comp += len(z)
        tr = orc.decode_trace(z, size)
        assert tr["status"] == 0 and tr["out"] == plain
        s = tr["seqs"]
        lls.append(s[:, 0]); mls.append(s[:, 1]); offs.append(s[:, 2])
        nseq_blocks += [int(x) for x in tr["block_nseq"] if x]
    ll, ml, off = np.concatenate(lls).astype(np.int64), np.concatenate(mls).astype(np.int64), np.concatenate(offs).astype(np.int64)
    n = len(off)
    out = dict(files=n_files, file_size=size, ratio=round(n_files * size / comp, 3), sequences=int(n),
               seq_per_block_mean=float(np.mean(nseq_blocks)),
               bytes_per_seq=float((ll.sum() + ml.sum()) / n), literal_fraction=float(ll.sum() / (ll.sum() + ml.sum())),
               ml_mean=float(ml.mean()), ll_mean=float(ll.mean()), ll_zero_fraction=float((ll == 0).mean()))
    out["ml_percentiles"] = {str(p): int(np.percentile(ml, p)) for p in (50, 75, 90, 95, 99, 99.9)}
    out["ll_percentiles"] = {str(p): int(np.percentile(ll, p)) for p in (50, 75, 90, 95, 99, 99.9)}
    out["ml_hist"] = {str(k): float(((ml >= k) & (ml < k2)).mean()) for k, k2 in ((3, 4), (4, 5), (5, 8), (8, 9), (9, 16), (16, 17), (17, 32), (32, 64), (64, 1 << 20))}
    edges = [1, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072, 262144, 524288, 1 << 20, 1 << 21]
    cum_seq = {str(e): float((off < e).mean()) for e in edges}
    cum_bytes = {str(e): float(ml[off < e].sum() / ml.sum()) for e in edges}
    out["offset_cdf_by_sequences"] = cum_seq
    out["offset_cdf_by_match_bytes"] = cum_bytes
    out["overlapping_fraction"] = float((off < ml).mean())
    print(json.dumps(out, indent=1))


main()
