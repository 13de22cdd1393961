"""Aggregate an ncu launch list (``--metrics gpu__time_duration.sum --csv``) per kernel (not a pytest file).

    python tools/summarize_launches.py launches.csv [steps] > summary.csv

`steps` = how many steps the profiled command ran (default 2: one warm-up + one timed); the per-step columns divide
by it.  ncu serialises kernels and starts each from a cold cache, so the SHARE of a kernel is what carries over to
the un-profiled step, not the absolute time."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        name = r[ki]
        m = re.match(r"(?:void )?((?:egp|at|c10)::(?:native::)?(?:<unnamed>::)?[A-Za-z0-9_]+)", name)
        key = m.group(1) if m else name.split("(")[0][:60]
        if key.startswith("egp::tc_gemm_kernel"):
            key += " CG2" if re.search(r", *\(int\)2>|, 2>", name.split("(const")[0]) else ""
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches_per_step", "us_per_step", "share_pct", "avg_us"])
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, round(n / steps, 1), round(us / steps, 1), round(100 * us / total, 2), round(us / n, 1)])
    w.writerow(["TOTAL", round(sum(a[0] for a in agg.values()) / steps, 1), round(total / steps, 1), 100.0, ""])


if __name__ == "__main__":
    main()
