#!/bin/bash
# all-reduce bucket size sweep at N GPUs (config 2): usage gpu_round2_p.sh N mb1 mb2 ...
mkdir -p gpurun_out
N=${1:-8}; shift
for mb in "$@"; do
  PE_B200_BUCKET_MB=$mb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --only-main --no-cpu-baseline --steps 20 --warmup 5 2> gpurun_out/r2p_${N}gpu_mb$mb.err | grep '^{' > gpurun_out/r2p_${N}gpu_mb$mb.json
  python - <<PY
import json
d = json.load(open("gpurun_out/r2p_${N}gpu_mb$mb.json"))
print("N=${N} bucket_mb=$mb", round(d["value"], 1), "samples/s", round(d["ms_per_step"], 3), "ms/step  e2e", round(d["e2e"]["value"], 1))
PY
done
