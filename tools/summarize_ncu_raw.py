"""Summarise `ncu --page raw --csv` exports (one kernel each) into a markdown table + profiles/<round>_traffic.json.

    python tools/summarize_ncu_raw.py gpurun_out/ncu_raw_*.csv             (round 1 names)
    ROUND=r2 python tools/summarize_ncu_raw.py gpurun_out/r2_ncu_raw_*.csv
"""
import csv
import json
import os
import sys

KEYS = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "DRAM read MB"), ("dram__bytes_write.sum", "DRAM write MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots % busy"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem KB")]


def to_mb(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)


def main(paths):
    out, traffic = [], {}
    for p in paths:
        rows = list(csv.reader(open(p)))
        if len(rows) < 3:
            continue
        hdr, units, r = rows[0], rows[1], rows[2]
        name = os.path.basename(p).replace("r2_ncu_raw_", "").replace("ncu_raw_", "").replace(".csv", "")
        kern = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("b200ocr::", "")
        vals = {}
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                v = r[i]
                if "bytes" in k:
                    v = "%.1f" % to_mb(v, units[i])
                elif k == "launch__shared_mem_per_block_dynamic":
                    v = "%.1f" % (to_mb(v, units[i]) * 1e3)
                else:
                    try:
                        v = "%.1f" % float(v.replace(",", ""))
                    except ValueError:
                        pass
                vals[label] = v
        out.append((name, kern, vals))
        traffic[name] = {"kernel": kern, "dram_read_mb": float(vals.get("DRAM read MB", 0)), "dram_write_mb": float(vals.get("DRAM write MB", 0)),
                         "time_us": float(vals.get("time us", 0))}
    labels = [l for _, l in KEYS]
    print("| capture | kernel | " + " | ".join(labels) + " |")
    print("|---|---|" + "---|" * len(labels))
    for name, kern, vals in out:
        print("| %s | `%s` | " % (name, kern[:48]) + " | ".join(vals.get(l, "") for l in labels) + " |")
    json.dump(traffic, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", os.environ.get("ROUND", "r1") + "_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1:])
