#!/usr/bin/env python
"""Markdown table (routine x dtype rows, order columns) of roofline fractions from tools/sweep_all_n.sh logs."""
import re
import sys

names = {"sym_solve": "`sym_solve`", "sym_matvec": "`sym_matvec`", "sym_invert": "`sym_invert`", "batch_inv": "`batchinv`",
         "batch_det": "`batchdet`", "batch_solve": "LU `solvevec`"}
cells = {}
for path in sys.argv[1:]:
    for l in open(path):
        m = re.match(r"(\w+) n=(\d+) (f32|f64)\s+[\d.]+ Gmat/s\s+[\d.]+ GB/s\s+frac ([\d.]+)", l)
        if m:
            cells[(m.group(1), m.group(3), int(m.group(2)))] = float(m.group(4))
print("| routine | dtype | n=1 | 2 | 3 | 4 | 5 | 6 | 7 | 8 | 9 | 10 |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for kind in ("sym_solve", "sym_matvec", "sym_invert", "batch_inv", "batch_det", "batch_solve"):
    for dt in ("f32", "f64"):
        row = [("%.2f" % cells[(kind, dt, n)]) if (kind, dt, n) in cells else "–" for n in range(1, 11)]
        print("| %s | %s | %s |" % (names[kind], dt, " | ".join(row)))
