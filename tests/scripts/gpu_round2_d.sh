#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=tests:rgb-proprioceptive-pose-estimator_b200:.
timeout 300 python - > gpurun_out/r2d_lstm.log 2>&1 <<'PY'
import torch, kernel_checks as kc
from pe_b200 import native
torch.backends.cuda.matmul.allow_tf32 = False
for fn in (lambda: kc.check_lstm_seq(3, 5, 64), lambda: kc.check_lstm_seq(20, 32, 512), lambda: kc.check_lstm_seq(10, 128, 512),
           lambda: kc.check_lstm_seq(4, 7, 512, with_state=True), lambda: kc.check_lstm_seq(2, 1, 512)):
    try:
        rows = fn(); torch.cuda.synchronize()
    except Exception as e:
        rows = [("EXCEPTION %r" % (e,), float("inf"), 0.0)]
    for n, e, t in rows:
        print("ok  " if e <= t else "FAIL", n, "%.3e" % e, flush=True)
    print("device flag", native.lib().pe_device_error(), flush=True)
PY
echo "lstm probe rc $?" >> gpurun_out/r2d_lstm.log; cat gpurun_out/r2d_lstm.log | tail -20
timeout 900 python tests/model_checks.py --forced tdo td tdo_v2 2>&1 | grep -v "^ok.*grad.*feature_net" > gpurun_out/r2d_forced.log; grep -c FAIL gpurun_out/r2d_forced.log; grep "worst\|median\|FAIL" gpurun_out/r2d_forced.log | cut -c1-120
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_dropin_scripts.py 2>&1 | tail -30 > gpurun_out/r2d_pytest.log; tail -4 gpurun_out/r2d_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2d_smoke.log 2>&1; tail -2 gpurun_out/r2d_smoke.log
timeout 600 python bench.py --model tdo --only-main --no-cpu-baseline > gpurun_out/r2d_bench_tdo.json 2> gpurun_out/r2d_bench_tdo.err; head -c 250 gpurun_out/r2d_bench_tdo.json; echo
timeout 600 python bench.py --only-main --no-cpu-baseline > gpurun_out/r2d_bench_no.json 2> gpurun_out/r2d_bench_no.err; head -c 250 gpurun_out/r2d_bench_no.json; echo
