#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line (instructions + stall samples)."""
import collections
import csv
import os
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
agg, samp = collections.Counter(), collections.Counter()
cur, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1], None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    try:
        agg[(cur, int(r[0]))] += int(d["Instructions Executed"])
        samp[(cur, int(r[0]))] += int(d["# Samples"])
    except (ValueError, KeyError):
        pass
tot, tots = sum(agg.values()), sum(samp.values())
print("total warp instructions %d, stall samples %d" % (tot, tots))
cache = {}
for (f, ln), v in sorted(agg.items(), key=lambda kv: -samp[kv[0]])[:top]:
    if f not in cache:
        try:
            cache[f] = open(f).read().splitlines()
        except OSError:
            cache[f] = []
    text = cache[f][ln - 1].strip()[:100] if ln - 1 < len(cache[f]) else ""
    print("%5.1f%% inst %5.1f%% samples  %s:%d  %s" % (100.0 * v / tot, 100.0 * samp[(f, ln)] / max(tots, 1), os.path.basename(f), ln, text))
