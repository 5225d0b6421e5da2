"""Per-instruction stall samples of one kernel of an ncu report (--page source --csv export), bucketed by address
range: where a kernel's time goes, and with how many active threads.   usage: ncu_buckets.py export.csv kernel [window]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
kern = sys.argv[2]; W = int(sys.argv[3]) if len(sys.argv) > 3 else 32
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
blk = [i for i in starts if kern in rows[i][1]][0]
end = min([i for i in starts if i > blk] + [len(rows)])
hdr = rows[blk + 1]; data = [r for r in rows[blk + 2:end] if len(r) == len(hdr)]
isrc = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed'); ithr = hdr.index('Avg. Threads Executed')
tot = sum(int(r[isamp]) for r in data); totex = sum(int(r[iex]) for r in data)
print('total samples', tot, 'warp instructions', totex, 'sass rows', len(data))
for i in range(0, len(data), W):
    s = sum(int(r[isamp]) for r in data[i:i + W]); ex = sum(int(r[iex]) for r in data[i:i + W])
    thr = [float(r[ithr]) for r in data[i:i + W] if int(r[iex]) > 0]
    if s > tot * 0.015:
        print('%5d-%5d samples %7d (%4.1f%%) exec %11d (%4.1f%%) avgthr %4.1f  %s' % (i, i + W, s, 100 * s / tot, ex, 100 * ex / totex, sum(thr) / max(1, len(thr)), data[i][isrc][:40]))
