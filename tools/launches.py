"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,...` launch list (development tool)."""
import csv
import sys

for f in sys.argv[1:]:
    print(f)
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    if not rows:
        print("  (no kernels)")
        continue
    hdr = rows[0]
    ik, im, iv, iid = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    d = {}
    for r in rows[1:]:
        d.setdefault((int(r[iid]), r[ik][:70]), {})[r[im]] = r[iv]
    for k, v in sorted(d.items()):
        print("  %3d %-70s %s" % (k[0], k[1], "  ".join("%s=%s" % (a.split('__')[1][:16], b) for a, b in v.items())))
