#!/usr/bin/env python
"""Summarise the SASS source page of an ncu report: instruction mix, top stall sites, phases.  usage: ncu_hot.py rep [kernel-index]"""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--print-details", "all"] if False else []), capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# split per kernel
kern = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kern.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
k = kern[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
h = {n: i for i, n in enumerate(k["hdr"])}
tot_inst = sum(int(r[h["Instructions Executed"]]) for r in k["rows"])
tot_samp = sum(int(r[h["# Samples"]]) for r in k["rows"])
print(k["name"][:90]); print("instructions executed", tot_inst, "samples", tot_samp)
mix = collections.Counter(); smp = collections.Counter()
for r in k["rows"]:
    op = r[h["Source"]].split()[0 if not r[h["Source"]].strip().startswith("@") else 1].split(".")[0]
    mix[op] += int(r[h["Instructions Executed"]]); smp[op] += int(r[h["# Samples"]])
print("opcode  inst%  sample%")
for op, c in mix.most_common(22):
    print(f"{op:10s} {100*c/tot_inst:5.1f} {100*smp[op]/max(1,tot_samp):5.1f}")
stalls = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.Counter()
for r in k["rows"]:
    for s in stalls:
        agg[s] += int(r[h[s]])
print("stall reasons:", [(s, round(100*v/max(1,sum(agg.values())),1)) for s, v in agg.most_common(8)])
print("top sample sites:")
for r in sorted(k["rows"], key=lambda r: -int(r[h["# Samples"]]))[:25]:
    top = max(stalls, key=lambda s: int(r[h[s]]))
    print(f'{r[h["Address"]][-5:]} {int(r[h["# Samples"]]):7d} {100*int(r[h["# Samples"]])/tot_samp:5.1f}%  {top:18s} {r[h["Source"]].strip()[:70]}')
