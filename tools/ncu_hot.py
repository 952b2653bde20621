"""Top stall-sample SASS instructions of one kernel in an .ncu-rep (ncu --page source --csv).
Usage: python tools/ncu_hot.py report.ncu-rep kernel-substring [topN]"""
import csv
import io
import subprocess
import sys


def main(rep, kern, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = [line]
            blocks.append(cur)
        elif cur is not None:
            cur.append(line)
    for b in blocks:
        if kern not in b[0]:
            continue
        rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
        hdr = rows[0]
        ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        data = rows[1:]
        total = sum(int(r[isamp]) for r in data)
        print(b[0][:120], "total samples", total)
        order = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top]
        for i in sorted(order):
            r = data[i]
            st = sorted(((int(r[c]), hdr[c]) for c in stall_cols), reverse=True)[:2]
            print(f"{i:5d} {int(r[isamp]):6d} {100.0 * int(r[isamp]) / total:5.1f}%  {r[isrc].strip()[:70]:70s} {st}")
        break


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
