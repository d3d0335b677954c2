#!/usr/bin/env python
"""Pivot a tune_main geometry sweep log: rows = geometry, columns = batch, cells = GB/s (best per column starred)."""
import re, sys
for path in sys.argv[1:]:
    rows, batches, first_op = {}, [], {}
    for l in open(path):
        m = re.match(r'(.+?)\s+batch\s+(\d+) T=\s*(\d+) thr=\s*(\d+) st=(\d) :\s+([\d.]+) us\s+([\d.]+) GB/s', l)
        if not m:
            continue
        op, b, T, thr, st, us, gb = m.groups()
        b = int(b)
        if b not in batches:
            batches.append(b)
        if op.strip() != first_op.setdefault(0, op.strip()):
            continue
        rows.setdefault((int(T), int(thr), int(st)), {})[b] = (float(us), float(gb))
    print(path)
    best = {b: max(v[b][1] for v in rows.values() if b in v) for b in batches}
    print("  %-22s" % "geometry" + "".join("%18d" % b for b in batches))
    for g, v in sorted(rows.items()):
        print("  T=%4d thr=%4d st=%d  " % g + "".join(
            ("%9.2fus %5.0f%s " % (v[b][0], v[b][1], "*" if v[b][1] >= best[b] * 0.995 else " ")) if b in v else " " * 18 for b in batches))
