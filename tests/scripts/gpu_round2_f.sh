#!/bin/bash
# GPU call F of round 2: ncu evidence for the final code (launch lists of one training step, --set full captures of
# the CTA-pair tap-GEMM, the narrow-tile tap-GEMM and the persistent LSTM kernels).  Every ncu command runs only after
# the same command exited 0 without ncu.
mkdir -p gpurun_out
export PYTHONPATH=rgb-proprioceptive-pose-estimator_b200:tests:.
bash tests/scripts/ncu_step_launches.sh r02 no 256
python tests/scripts/launch_list_summary.py gpurun_out/launches_r02.csv gpurun_out/launches_r02_train_step_no_b256_summary.json "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python tests/profile_step.py no 256 1" | tail -16
bash tests/scripts/ncu_step_launches.sh r02_tdo tdo 32
python tests/scripts/launch_list_summary.py gpurun_out/launches_r02_tdo.csv gpurun_out/launches_r02_train_step_tdo_n32_s20_summary.json "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python tests/profile_step.py tdo 32 1" | tail -8
bash tests/scripts/ncu_conv.sh r02_fwd3x3_256at14_pair 256 14 256 256 3 1 fwd 1
bash tests/scripts/ncu_conv.sh r02_fwd1x1_256to1024at14_pair 256 14 256 1024 1 1 fwd 1
bash tests/scripts/ncu_conv.sh r02_dgrad1x1_2048to512at7_pair 256 7 2048 512 1 1 dgrad
bash tests/scripts/ncu_conv.sh r02_fwd3x3_64at56 256 56 64 64 3 1 fwd 1
bash tests/scripts/ncu_conv.sh r02_fwd3x3_128at28 256 28 128 128 3 1 fwd 1
bash tests/scripts/ncu_conv.sh r02_fwd1x1_64to256at56 256 56 64 256 1 1 fwd 1
bash tests/scripts/ncu_conv.sh r02_wgrad3x3_256at14 256 14 256 256 3 1 wgrad
# persistent LSTM kernels inside one TDO step
CMD="python tests/profile_step.py tdo 32 1"
ncu --set full --clock-control none --import-source on -k regex:lstm_seq -c 2 -f -o gpurun_out/prof_r02_lstm_seq $CMD > gpurun_out/ncu_r02_lstm_seq.log 2>&1; echo "lstm rc=$?"
python tests/scripts/ncu_summary.py gpurun_out/conv_r02_summary.csv gpurun_out/prof_r02_*.ncu-rep
ls -la gpurun_out/*.ncu-rep | awk '{s+=$5} END {print s/1e6, "MB of reports"}'
rm -f gpurun_out/prof_r02_fwd3x3_128at28.ncu-rep gpurun_out/prof_r02_fwd1x1_64to256at56.ncu-rep
