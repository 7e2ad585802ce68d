#!/bin/bash
# exposed share of the gradient all-reduce at N GPUs: config-2 step with and without the collective
mkdir -p gpurun_out
N=${1:-8}
for skip in 0 1; do
  PE_B200_SKIP_ALLREDUCE=$skip timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --only-main --no-cpu-baseline --steps 20 --warmup 5 2> gpurun_out/r2o_${N}gpu_skip$skip.err | grep '^{' > gpurun_out/r2o_${N}gpu_skip$skip.json
  python - <<PY
import json
d = json.load(open("gpurun_out/r2o_${N}gpu_skip$skip.json"))
print("N=${N} skip_allreduce=$skip", round(d["value"], 1), "samples/s", round(d["ms_per_step"], 3), "ms/step  e2e", round(d["e2e"]["value"], 1))
PY
done
