#!/bin/bash
# ncu evidence of the round-2b code: launch list of one training step + full captures of the narrow-tile launches
mkdir -p gpurun_out
export PYTHONPATH=rgb-proprioceptive-pose-estimator_b200:tests:.
# usage: TAG_PREFIX=r02c bash tests/scripts/gpu_round2_m.sh
bash tests/scripts/ncu_step_launches.sh ${TAG_PREFIX:-r02b}_train_step_no_b256 no 256
cap() {   # cap <tag> <script> <args...>
    TAG=$1; shift
    python "$@" > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed: $TAG"; return; }
    ncu --set full --clock-control none --import-source on -k regex:tapgemm -s 1 -c 1 -f -o gpurun_out/prof_$TAG \
        python "$@" > gpurun_out/ncu_$TAG.log 2>&1
    echo "$TAG rc=$?"
}
cap ${TAG_PREFIX:-r02b}_stem_fwd_s2d tests/profile_stem.py 256 fwd
cap ${TAG_PREFIX:-r02b}_stem_wgrad_s2d tests/profile_stem.py 256 wgrad
cap ${TAG_PREFIX:-r02b}_fwd3x3_64at56_pair tests/profile_conv.py 256 56 64 64 3 1 fwd 1
cap ${TAG_PREFIX:-r02b}_fwd3x3_128at28_pair tests/profile_conv.py 256 28 128 128 3 1 fwd 1
cap ${TAG_PREFIX:-r02b}_fwd1x1_64to64at56_pair tests/profile_conv.py 256 56 64 64 1 1 fwd 1
ls -la gpurun_out/*.ncu-rep | tail
