#!/usr/bin/env python
"""Condense `ncu --set full` reports into the few metrics the roofline argument needs.
usage: ncu_summary.py out.csv report1.ncu-rep [report2.ncu-rep ...]   (runs `ncu -i ... --page raw --csv`)"""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__cluster_dim_x", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]
SCALE_T = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
SCALE_B = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    rows_out = []
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
        for r in rows[2:]:
            d = {"report": rep.split("/")[-1]}
            for w, i in idx:
                d[w + (" [%s]" % units[i] if units[i] else "")] = r[i]
            try:
                it = hdr.index("gpu__time_duration.sum")
                t_s = float(r[it]) * SCALE_T.get(units[it], 1e-6)
                b = 0.0
                for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    i = hdr.index(k)
                    b += float(r[i]) * SCALE_B.get(units[i], 1)
                d["dram_GB_per_s"] = "%.1f" % (b / t_s / 1e9)
            except Exception:
                pass
            rows_out.append(d)
    keys = []
    for d in rows_out:
        for k in d:
            if k not in keys:
                keys.append(k)
    with open(out, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        for d in rows_out:
            w.writerow(d)
    print("wrote %d rows to %s" % (len(rows_out), out))


if __name__ == "__main__":
    main()
