"""Run-to-run spread of the 100-step TDO loss curve (atomic summation order) vs the reference fixture."""
import sys
import torch
import curve_check as cc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for r in range(n):
    losses, ref, dev = cc.run_curve("tdo", fixture="curve_tdo_lr1e-5.json")
    lag = []
    for i in range(10, 25):
        lo, hi = min(ref[i - 2:i + 3]), max(ref[i - 2:i + 3])
        v = losses[i]
        lag.append(0.0 if lo <= v <= hi else min(abs(v - lo) / lo, abs(v - hi) / hi))
    pm = sum(ref[25:]) / len(ref[25:])
    print("run %d: first10 %.3f  steep(10-25) %.3f  lag-tolerant %.3f  plateau min/max ratio %.2f %.2f  tail ratio %.3f"
          % (r, max(dev[:10]), max(dev[10:25]), max(lag), min(losses[25:]) / pm, max(losses[25:]) / pm,
             (sum(losses[-30:]) / 30) / (sum(ref[-30:]) / 30)), flush=True)
