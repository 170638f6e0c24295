"""Summarise an ncu launch list (``ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <command>``) into the
per-kernel table kept under profiles/: launches, total and average device time, share of the captured time, and the class shares
(GEMM / attention / row-wise / conditioning) that bench.py's CUDA-event profile reports for the same step.
usage: python tools/ncu_launch_list.py <launches.csv> <out.txt> [title]"""
import csv
import re
import sys
from collections import OrderedDict


def short_name(k: str) -> str:
    k = re.sub(r"^void\s+", "", k)
    k = re.sub(r"^ndit::", "", k)
    k = re.sub(r"\(.*$", "", k)            # drop the parameter list
    return k.replace("ndit::", "")


def klass(name: str) -> str:
    if name.startswith("gemm"):
        return "GEMM"
    if name.startswith("attention"):
        return "attention"
    if name.startswith(("gemv_rows", "cond_prepare", "moe_time_select", "gather_label")):
        return "conditioning"
    return "row-wise"


def main():
    src, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else src
    rows = []
    with open(src, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]          # skip ==PROF== chatter and the program's own output
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        us = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(unit, 1e-3)
        rows.append((short_name(r[ik]), us))
    agg = OrderedDict()
    for n, us in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    cls = {}
    for n, (c, t) in agg.items():
        cls[klass(n)] = cls.get(klass(n), 0.0) + t
    with open(out, "w") as f:
        f.write(f"# {title}\n# ncu --metrics gpu__time_duration.sum --clock-control none: device time per launch (cold-cache, serialised, profiler clocks:\n"
                "# compare SHARES with bench.py's CUDA-event profile of the same step, not absolutes)\n")
        f.write(f"# launches {len(rows)}   total {total / 1e3:.3f} ms   " +
                "  ".join(f"{k} {100 * v / total:.1f} %" for k, v in sorted(cls.items(), key=lambda kv: -kv[1])) + "\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{n:<64} launches {c:5d}  total {t:12.1f} us  avg {t / c:9.1f} us  share {100 * t / total:5.1f} %\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
