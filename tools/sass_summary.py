"""Per-kernel SASS evidence of the Blackwell-native path (B200_PROFILING.md "What proves a Blackwell-native kernel"):
counts of UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UBLKCP (TMA), LDGSTS (cp.async),
HMMA (legacy mma.sync: must be 0) per kernel of tsasr_b200/libtsasr_b200.so.  python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tsasr_b200", "libtsasr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "LDGSTS", "MUFU.EX2", "HMMA", "HGMMA"]
kernels, cur = {}, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        kernels[cur] = {p: 0 for p in pats}
        kernels[cur]["instructions"] = 0
        continue
    if cur and re.search(r"/\*[0-9a-f]{4,}\*/", line):
        kernels[cur]["instructions"] += 1
        for p in pats:
            if re.search(r"\b" + re.escape(p) + (r"\b" if "." in p else r"(\b|\.)"), line):
                if p == "UTCHMMA" or p != "UTCHMMA.2CTA" or ".2CTA" in line:
                    kernels[cur][p] += 1
arch = re.findall(r"arch = (sm_\w+)", out)
print(f"cuobjdump -sass {os.path.relpath(so, ROOT)}   (architectures: {sorted(set(arch))})")
print(f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{p:>12s}" for p in pats))
for k, c in sorted(kernels.items()):
    print(f"{k[:58]:58s} {c['instructions']:6d} " + " ".join(f"{c[p]:12d}" for p in pats))
tot = {p: sum(c[p] for c in kernels.values()) for p in pats}
print(f"{'total':58s} {sum(c['instructions'] for c in kernels.values()):6d} " + " ".join(f"{tot[p]:12d}" for p in pats))
