"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time and share of
one bench step (from one forward joint GEMM launch to the next).  Usage:
    python tools/launch_summary.py gpurun_out/launches.csv [step_index]      (default: the last step of the value loop)"""
import collections
import csv
import re
import sys


def main(path, which=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [(r["Kernel Name"], float(r["Metric Value"])) for r in rows]
    is_fwd = re.compile(r"joint_gemm_kernel<\(?(tsasr::)?(JointMode\))?0")
    fwd = [i for i, (n, _) in enumerate(names) if is_fwd.search(n)]
    if not fwd:
        raise SystemExit("no forward joint GEMM launch in the list")
    # steps of the device-timed loop are back to back (flush + our kernels only); the e2e loop adds torch kernels
    steps = []
    for a, b in zip(fwd, fwd[1:] + [len(names)]):
        steps.append((a, b))
    if which is None:
        sizes = [b - a for a, b in steps]
        small = min(sizes)
        which = max(i for i, s in enumerate(sizes) if s == small)  # last step of the shortest kind
    start, end = steps[which]
    agg = collections.OrderedDict()
    for n, v in names[start:end]:
        k = re.sub(r"\(.*", "", n)[:72]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"step {which} of {len(steps)} (launches {start}..{end - 1})")
    print(f"{'kernel':74s} {'n':>4s} {'total ms':>9s} {'avg us':>9s} {'share':>6s}")
    for k, a in agg.items():
        print(f"{k:74s} {a[0]:4d} {a[1] / 1e6:9.3f} {a[1] / a[0] / 1e3:9.1f} {a[1] / tot * 100:5.1f}%")
    print(f"{'total':74s} {'':4s} {tot / 1e6:9.3f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
