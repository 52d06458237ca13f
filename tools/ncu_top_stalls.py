"""Top stalled SASS instructions from `ncu -i rep --page source --csv` output (file path as argv[1])."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Source' in r][0]
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[hi + 1:]:
    try:
        n = int(r[ci['# Samples']])
    except Exception:
        continue
    top = sorted(((int(r[ci[s]] or 0), s) for s in stalls), reverse=True)[:2]
    data.append((n, r[ci['Source']][:95], top, r[ci['Instructions Executed']]))
tot = sum(d[0] for d in data)
print("total samples", tot)
for n, src, top, ex in sorted(data, key=lambda x: -x[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%6d %5.1f%% ex=%-8s %-95s %s" % (n, 100.0 * n / max(tot, 1), ex, src, " ".join("%s=%d" % (s[6:], c) for c, s in top if c)))
