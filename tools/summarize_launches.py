"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches, total and share."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
    agg = defaultdict(lambda: [0, 0.0])
    for k, ns in rows:
        name = re.sub(r"\(.*", "", k).replace("void ", "").replace("b200ocr::", "")
        agg[name][0] += 1
        agg[name][1] += ns
    total = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | share |")
    print("|---|---|---|---|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, n, ns / 1e3, 100 * ns / total))
    print("\n%d launches, %.3f ms total (cold-cache, serialised: compare shares, not absolutes)" % (len(rows), total / 1e6))


if __name__ == "__main__":
    main(sys.argv[1])
