#!/usr/bin/env python
"""Aggregate an ncu launch list (tests/scripts/ncu_step_launches.sh) per kernel: launches, time share, DRAM bytes.
usage: launch_list_summary.py launches.csv out.json "<command that produced it>" """
import collections
import csv
import json
import sys

T = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
B = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    lines = [l for l in open(src) if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    iK, iM, iV, iU, iID = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = collections.OrderedDict()
    for r in rd:
        d = per.setdefault(r[iID], {"k": r[iK].split("(")[0].split("::")[-1]})
        v = float(r[iV].replace(",", ""))
        if r[iM] == "gpu__time_duration.sum":
            d["us"] = v * T.get(r[iU], 1.0)
        else:
            d[r[iM]] = v * B.get(r[iU], 1)
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["k"], {"launches": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["launches"] += 1
        a["us"] += d.get("us", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a["us"] for a in agg.values())
    kernels = []
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        kernels.append({"kernel": k, "launches": a["launches"], "ms": round(a["us"] / 1e3, 3),
                        "share": round(a["us"] / tot, 4), "dram_read_GB": round(a["rd"] / 1e9, 3),
                        "dram_write_GB": round(a["wr"] / 1e9, 3),
                        "dram_GB_per_s": round((a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9, 1) if a["us"] else None})
    json.dump({"command": cmd, "note": "ncu per-launch times are cold-cache and serialised: compare shares, not absolutes",
               "total_ms": round(tot / 1e3, 3), "launches": sum(a["launches"] for a in agg.values()), "kernels": kernels},
              open(out, "w"), indent=1)
    for k in kernels[:12]:
        print(k)
    print("total ms %.2f, launches %d" % (tot / 1e3, sum(a["launches"] for a in agg.values())))


if __name__ == "__main__":
    main()
