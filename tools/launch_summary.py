"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time and share of
the LAST bench step (between the last two forward joint GEMM launches).  Usage:
    python tools/launch_summary.py gpurun_out/launches.csv [nth_fwd_from_end]"""
import collections
import csv
import sys


def main(path, which=-1):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [(r["Kernel Name"], float(r["Metric Value"])) for r in rows]
    fwd = [i for i, (n, _) in enumerate(names) if "joint_gemm_kernel<0>" in n]
    # the value-loop steps come first; the e2e loop follows (it has extra torch kernels)
    start = fwd[which]
    end = fwd[which + 1] if which + 1 < 0 and which + 1 != 0 else len(names)
    if which != -1:
        end = fwd[which + 1]
    agg = collections.OrderedDict()
    for n, v in names[start:end]:
        k = n.split("(")[0][:70]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':72s} {'n':>4s} {'total ms':>9s} {'avg us':>9s} {'share':>6s}")
    for k, a in agg.items():
        print(f"{k:72s} {a[0]:4d} {a[1] / 1e6:9.3f} {a[1] / a[0] / 1e3:9.1f} {a[1] / tot * 100:5.1f}%")
    print(f"{'total':72s} {'':4s} {tot / 1e6:9.3f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else -1)
