"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) of bench.py into the
per-kernel shares of ONE steady-state step (the launches between the last two `axpbz_kernel` launches).

    python tools/launch_list.py gpurun_out/launches.csv > profiles/rNN_launch_list.md
"""

import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], us))
    marks = [i for i, (k, _) in enumerate(rows) if "axpbz_kernel" in k]
    if len(marks) < 2:
        raise SystemExit("need at least two purification steps in the capture")
    step = rows[marks[-2]:marks[-1]]
    agg = OrderedDict()
    for k, us in step:
        name = k.split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in agg.values())
    print("step total: %.2f ms over %d launches\n" % (total / 1e3, len(step)))
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (name[:80], n, us, 100 * us / total))
    ours = sum(us for name, (n, us) in agg.items() if "ap::" in name or name.startswith("ap"))
    print("\nour kernels (`ap::*`): %.1f%% of the step" % (100 * ours / total))


if __name__ == "__main__":
    main()
