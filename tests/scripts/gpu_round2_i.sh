#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=tests:rgb-proprioceptive-pose-estimator_b200:.
timeout 300 python - > gpurun_out/r2i_dgradbn.log 2>&1 <<'PY'
import torch, kernel_checks as kc
from pe_b200 import native
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
for fn in (lambda: kc.check_conv_dgrad_bn(3, 28, 28, 128, 512, 1, 1), lambda: kc.check_conv_dgrad_bn(2, 56, 56, 64, 64, 3, 1),
           lambda: kc.check_conv_dgrad_bn(5, 14, 14, 256, 1024, 1, 1), lambda: kc.check_conv_dgrad_bn(3, 14, 14, 256, 256, 3, 1),
           lambda: kc.check_conv_dgrad_bn(2, 28, 28, 256, 256, 3, 2), lambda: kc.check_conv_dgrad_bn(3, 7, 7, 512, 2048, 1, 1),
           lambda: kc.check_conv_dgrad_bn(64, 56, 56, 64, 256, 1, 1), lambda: kc.check_lstm_seq(20, 32, 512)):
    try:
        rows = fn(); torch.cuda.synchronize()
    except Exception as e:
        rows = [("EXCEPTION %r" % (e,), float("inf"), 0.0)]
    for n, e, t in rows:
        print("ok  " if e <= t else "FAIL", n, "%.3e" % e, flush=True)
    f = native.lib().pe_device_error()
    if f:
        print("device flag", f, flush=True); native.lib().pe_device_error_clear()
PY
cat gpurun_out/r2i_dgradbn.log | tail -20
timeout 900 python tests/model_checks.py --forced no tdo n=4 2>&1 | grep "worst\|median\|FAIL\|forward" | cut -c1-120 > gpurun_out/r2i_forced.log; cat gpurun_out/r2i_forced.log
for f in 1 0; do
PE_FUSE=$f timeout 600 python - <<'PY' 2>&1 | tail -3
import os, sys, json, subprocess
sys.path.insert(0, "rgb-proprioceptive-pose-estimator_b200")
from pe_b200 import engine
engine.FUSE_BN_REDUCE[0] = os.environ["PE_FUSE"] == "1"
sys.argv = ["bench.py", "--only-main", "--no-cpu-baseline", "--steps", "10"]
import runpy
try:
    runpy.run_path("bench.py", run_name="__main__")
except SystemExit:
    pass
PY
done > gpurun_out/r2i_ab.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2i_ab.log"):
    if l.startswith("{"):
        d = json.loads(l); k = d["kernel_ms"]
        print(round(d["value"], 1), round(d["ms_per_step"], 3), {n: k.get(n) for n in ("pe_conv2d_dgrad", "pe_conv2d_dgrad_bn", "pe_bn_bwd_reduce", "pe_bn_bwd_apply")})
PY
