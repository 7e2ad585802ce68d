#!/bin/bash
# 2-GPU call: persistent-LSTM kernel checks, two-GPU parity tests, bench at N = 2
mkdir -p gpurun_out
export PYTHONPATH=tests:rgb-proprioceptive-pose-estimator_b200:.
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "36 or 37 or 35" 2>&1 | tail -5 > gpurun_out/r2e_lstm.log; tail -3 gpurun_out/r2e_lstm.log
timeout 1500 python -m pytest tests/test_ddp_gpu.py -q -s 2>&1 | tail -60 > gpurun_out/r2e_ddp.log; cat gpurun_out/r2e_ddp.log | tail -14
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err
head -c 300 gpurun_out/r2e_bench_2gpu.json; echo; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2e_bench_2gpu.json"))
print("no 2gpu", d["value"], d["e2e"]["value"], "tdo", d["configs"]["tdo"]["value"], d["configs"]["tdo"]["e2e"]["value"], d["configs"]["tdo"]["e2e_u8"]["value"])
PY
