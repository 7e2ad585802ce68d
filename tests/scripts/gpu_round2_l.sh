#!/bin/bash
# 8-GPU call: the driver's scaling command at N = 8 (config 2 as `value`, config 4 as configs.tdo)
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/r2l_bench_${N}gpu.err | grep '^{' > gpurun_out/r2l_bench_${N}gpu.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2l_bench_${N}gpu.json"))
t = d["configs"]["tdo"]
print("N=${N} no", round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "| tdo", round(t["value"], 1), round(t["ms_per_step"], 3), "e2e", round(t["e2e"]["value"], 1), "u8", round(t["e2e_u8"]["value"], 1))
PY
tail -3 gpurun_out/r2l_bench_${N}gpu.err
