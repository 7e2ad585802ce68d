#!/bin/bash
# GPU call A of round 2: full GPU test suite, teacher-forced gradient table, lr 1e-3 curve, bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt; free -g >> gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_dropin_scripts.py 2>&1 | tail -40 > gpurun_out/r2a_pytest.log
timeout 900 python -m pytest tests/test_dropin_scripts.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2a_dropin.log
timeout 900 python tests/model_checks.py --forced 2>&1 | grep -v "^ok.*grad.*feature_net" > gpurun_out/r2a_forced.log
timeout 600 python tests/curve_lr1e3.py > gpurun_out/r2a_curve.log 2>&1
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_pytest.log; tail -3 gpurun_out/r2a_dropin.log; tail -3 gpurun_out/r2a_curve.log; head -c 600 gpurun_out/r2a_bench.json
