#!/bin/bash
# 2-GPU sweep: NCCL CTA cap x SMs reserved by the persistent GEMM grids during the backward pass
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --only-main --model ${MODEL:-no} 2> gpurun_out/r2g_$name.err | grep '^{' > gpurun_out/r2g_$name.json
  python -c "
import json,sys; d=json.load(open('gpurun_out/r2g_$name.json')); print('$name', '${MODEL:-no}', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
}
run base A=1
run res8 PE_B200_SM_RESERVE=8
run res16 PE_B200_SM_RESERVE=16
run cta4 NCCL_MAX_CTAS=4
run cta8_res8 NCCL_MAX_CTAS=8 PE_B200_SM_RESERVE=8
run cta4_res4 NCCL_MAX_CTAS=4 PE_B200_SM_RESERVE=4
MODEL=tdo run tdo_base A=1
MODEL=tdo run tdo_cta4_res4 NCCL_MAX_CTAS=4 PE_B200_SM_RESERVE=4
