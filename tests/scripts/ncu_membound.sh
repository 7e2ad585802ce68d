#!/bin/bash
# ncu --set full captures of the HBM-bound kernels of one training step (profiles/ recipe, run under gpurun).
# usage: tests/scripts/ncu_membound.sh <tag>
set -u
TAG=${1:-r01}
export PYTHONPATH=rgb-proprioceptive-pose-estimator_b200:tests:.
CMD="python tests/profile_step.py no 256 1"
$CMD > gpurun_out/plain_step_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_step_$TAG.log; exit 1; }
for K in bn_apply_kernel bn_bwd_apply_kernel channel_reduce_kernel adam_kernel im2col_stem_kernel maxpool_fwd_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -c 2 -f -o gpurun_out/prof_${K}_$TAG $CMD \
      > gpurun_out/ncu_${K}_$TAG.log 2>&1
  echo "$K rc=$?"
done
