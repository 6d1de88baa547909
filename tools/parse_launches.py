#!/usr/bin/env python3
"""Pair the LAUNCHED lines of tools/ncu_variants.py with the ncu --csv launch list."""
import csv, json, sys
order = [json.loads(l.split(" ", 1)[1]) for l in open(sys.argv[1]) if l.startswith("LAUNCHED")]
rows = [r for r in csv.reader(l for l in open(sys.argv[2]) if not l.startswith("==")) if len(r) > 5]
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
main = [float(r[vi].replace(",", "")) for r in rows if "haar_icon_tma" in r[ki]]
strip = [float(r[vi].replace(",", "")) for r in rows if "edge_strip" in r[ki]]
for i, o in enumerate(order):
    m = main[2 * i: 2 * i + 2]; s = strip[2 * i: 2 * i + 2]
    print(o["depths"], o["variant"], "main_us", [round(x / 1e3, 1) for x in m], "strip_us", [round(x / 1e3, 1) for x in s])
