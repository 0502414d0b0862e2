#!/usr/bin/env python
"""Turn what `scripts/gpu_round.sh rNN` left in gpurun_out/ into the tracked evidence under profiles/ (run here, no GPU):
bench lines copied as they are, every .ncu-rep summarised (scripts/ncu_summary.py), the launch list reduced to per-kernel
shares, and profiles/traffic.json refreshed from the full captures (the per-launch DRAM bytes bench.py quotes as STATIC).
usage: make_profiles.py r02"""
import collections
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

for f in sorted(glob.glob(os.path.join(G, R + "_bench_*.json")) + glob.glob(os.path.join(G, R + "_detransform.json"))):
    txt = open(f).read().strip()
    if not txt:
        continue
    line = txt.splitlines()[-1]
    try:
        json.loads(line)
    except ValueError:
        continue
    open(os.path.join(P, os.path.basename(f)), "w").write(line + "\n")
    print("bench line", os.path.basename(f))

traffic = {}


def num(s):
    v, _, u = s.partition(" ")
    return float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u.strip(), 1.0)


# gpu_round.sh summarises every .ncu-rep on the box (gpurun copies back at most 64 MiB): <R>_<name>.ncu.json holds one
# entry per captured kernel, <R>_<name>.by_line.txt the top source lines; a report that did come back is summarised here
for rep in sorted(glob.glob(os.path.join(G, "prof_" + R + "_*.ncu-rep"))):
    name = os.path.basename(rep)[len("prof_"):-len(".ncu-rep")]
    if not os.path.exists(os.path.join(G, name + ".ncu.json")):
        subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, os.path.join(G, name + ".ncu.json")], capture_output=True)
for f in sorted(glob.glob(os.path.join(G, R + "_*.ncu.json"))):
    name = os.path.basename(f)[:-len(".ncu.json")]
    try:
        ds = json.load(open(f))
    except ValueError:
        print("could not read", f)
        continue
    shutil.copy(f, os.path.join(P, name + ".ncu.json"))
    for d in ds:
        kern = d.get("kernel", "").split("(")[0].split("<")[0].replace("_kernel", "").replace("void ", "").strip()
        if "dram_read" in d and "dram_write" in d:
            # the bench configuration the capture was taken at (scripts/gpu_round.sh); bench.py quotes the bytes only for that one
            at = {"zstd_enc_blocks": "zstd+aes|speed|K|1024", "gcm_main": "zstd+aes|speed|K|1024",
                  "zstd_enc_regions": "zstd+aes|dense|K|1024"}.get(kern)
            traffic[kern] = {"bytes": num(d["dram_read"]) + num(d["dram_write"]), "capture": "profiles/%s.ncu.json" % name,
                             "duration_under_ncu": d.get("duration"), "bench_config": at}
        print("ncu summary", name, kern, d.get("duration"), d.get("issue_active_pct"))
if traffic:
    tpath = os.path.join(P, "traffic.json")
    merged = {}
    if os.path.exists(tpath):                     # keep the kernels this run did not capture (a later partial run must not drop them)
        try:
            merged = json.load(open(tpath))
        except ValueError:
            merged = {}
    merged.update(traffic)
    json.dump(merged, open(tpath, "w"), indent=1)
for f in sorted(glob.glob(os.path.join(G, R + "_*.by_line.txt")) + glob.glob(os.path.join(G, R + "_sanitizer_*")) +
                glob.glob(os.path.join(G, R + "_pcie_ceiling*.json"))):
    if os.path.getsize(f) < (1 << 20):
        shutil.copy(f, os.path.join(P, os.path.basename(f)))
        print("copied", os.path.basename(f))

lst = os.path.join(G, R + "_launches.csv")
if os.path.exists(lst):
    shutil.copy(lst, os.path.join(P, R + "_launches_zstdaes_1GiB.csv"))
    rows = [r for r in csv.reader(open(lst)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    acc = collections.Counter(); cnt = collections.Counter()
    for r in rows:
        if r is hdr or len(r) <= vi:
            continue
        try:
            acc[r[ki].split("(")[0]] += float(r[vi].replace(",", "")); cnt[r[ki].split("(")[0]] += 1
        except ValueError:
            pass
    tot = sum(acc.values())
    json.dump({k: {"launches": cnt[k], "ns_total": acc[k], "share": acc[k] / tot} for k in acc}, open(os.path.join(P, R + "_launch_shares_zstdaes_1GiB.json"), "w"), indent=1)
    print("launch shares", {k: round(v / tot, 3) for k, v in acc.items()})
