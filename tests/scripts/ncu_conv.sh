#!/bin/bash
# ncu --set full capture of one tap-GEMM configuration (profiles/ recipe, run under gpurun).
# usage: tests/scripts/ncu_conv.sh <tag> <profile_conv.py args...>
set -u
TAG=$1; shift
export PYTHONPATH=rgb-proprioceptive-pose-estimator_b200:tests:.
python tests/profile_conv.py "$@" > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:tapgemm -s 1 -c 1 -f -o gpurun_out/prof_$TAG \
    python tests/profile_conv.py "$@" > gpurun_out/ncu_$TAG.log 2>&1
echo "$TAG rc=$?"
