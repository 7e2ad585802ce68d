#!/bin/bash
# Every launch of one training step with its device time and DRAM bytes (profiles/ recipe, run under gpurun).
# usage: tests/scripts/ncu_step_launches.sh <tag> [model] [batch]
set -u
TAG=${1:-r01}; MODEL=${2:-no}; BATCH=${3:-256}
export PYTHONPATH=rgb-proprioceptive-pose-estimator_b200:tests:.
CMD="python tests/profile_step.py $MODEL $BATCH 1"
$CMD > gpurun_out/plain_step_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_step_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
