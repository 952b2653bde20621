"""Warp-stall sampling summary of one kernel from an ncu report (evidence tool, runs without a GPU).

    ncu -i rep.ncu-rep --page source --csv --kernel-id ::regex:<name>:<n> > k.csv ; python tools/stall_summary.py k.csv

Prints the share of each stall reason over all warp samples, and the same per 2 KB region of SASS with the instruction
kinds that identify the role living there (MUFU/LDTM: epilogue, LDG/F2FP: A producers, SYNCS: barrier polling,
UCGABAR: the cluster barrier where role-less warps park for the whole kernel)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[hi[0]]
data = [r for r in rows[hi[0] + 1: (hi[1] if len(hi) > 1 else len(rows))] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(rows[0][1] if len(rows[0]) > 1 else "")
tot = collections.Counter()
for r in data:
    for s in stalls:
        tot[s[6:]] += int(r[col[s]] or 0)
S = sum(tot.values())
print(f"warp samples: {S}")
print("  " + "  ".join(f"{k} {100 * v / S:.1f}%" for k, v in tot.most_common(10)))
base = int(data[0][0], 16)
reg = collections.OrderedDict()
for r in data:
    k = (int(r[0], 16) - base) // 0x800
    d = reg.setdefault(k, {"n": 0, "st": collections.Counter(), "kinds": collections.Counter()})
    d["n"] += int(r[col["# Samples"]] or 0)
    for s in stalls:
        d["st"][s[6:]] += int(r[col[s]] or 0)
    src = r[col["Source"]]
    for pat in ("MUFU", "LDTM", "UTC", "STG", "LDG", "F2FP", "SYNCS", "UCGABAR", "BAR.SYNC"):
        if pat in src:
            d["kinds"][pat] += 1
print("per 2 KB SASS region (regions with >= 1% of the samples):")
for k, d in reg.items():
    if d["n"] >= 0.01 * S:
        kinds = " ".join(f"{a}:{b}" for a, b in d["kinds"].most_common(4))
        top = "  ".join(f"{a} {100 * b / d['n']:.0f}%" for a, b in d["st"].most_common(4))
        print(f"  +0x{k * 0x800:05x}  {100 * d['n'] / S:5.1f}% of samples   [{kinds:38s}]  {top}")
