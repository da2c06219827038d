"""Top stall sites from an `ncu --page source --csv` dump (development tool)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
isrc, isamp, iex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
ilsb = hdr.index('stall_long_sb')
data = []
for k, r in enumerate(rows[2:]):
    try:
        data.append((float(r[isamp]), k, r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print('instructions', len(data), 'total samples', tot)
for v, k, r in sorted(data, key=lambda t: -t[0])[:top]:
    print(f"{k:5d} {v:8.0f} {100 * v / tot:5.1f}%  long_sb={r[ilsb]:>6s} exec={r[iex]:>9s}  {r[isrc].strip()[:90]}")
