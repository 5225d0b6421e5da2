"""Per-segment instruction accounting of one kernel from an `ncu --page source --csv` export: consecutive SASS rows with
similar execution counts are merged; counts are shown per `unit` executions of the kernel's outermost loop (e.g. per
warp-round).   usage: ncu_segments.py export.csv unit_count [min_per_unit]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
R = float(sys.argv[2]); thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 8
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
isrc = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed'); ithr = hdr.index('Avg. Threads Executed')
seg = []; cur = None
for i, r in enumerate(data):
    ex = int(r[iex]); thr = float(r[ithr]); smp = int(r[isamp])
    if cur and ex > 0 and abs(ex - cur['e0']) <= 0.3 * cur['e0']:
        cur['n'] += 1; cur['ex'] += ex; cur['smp'] += smp; cur['thr'] += thr * ex; cur['last'] = i
    else:
        if cur: seg.append(cur)
        cur = dict(first=i, last=i, n=1, e0=max(ex, 1), ex=ex, smp=smp, thr=thr * ex, src=r[isrc].strip()[:50])
seg.append(cur)
tot = sum(s['ex'] for s in seg); ts = sum(s['smp'] for s in seg)
for s in seg:
    if s['ex'] / R >= thresh or s['smp'] > 0.01 * ts:
        print('%4d-%4d n%3d  per-unit %7.1f  x%6.2f thr %4.1f smp %4.1f%%  %s' % (s['first'], s['last'], s['n'], s['ex'] / R, s['e0'] / R, s['thr'] / max(s['ex'], 1), 100 * s['smp'] / ts, s['src']))
print('total per unit %.1f, rows %d' % (tot / R, len(data)))
