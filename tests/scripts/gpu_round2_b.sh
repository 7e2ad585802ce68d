#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python tests/model_checks.py --forced 2>&1 | grep -v "^ok.*grad.*feature_net" > gpurun_out/r2b_forced.log)
(timeout 600 python tests/model_checks.py --forced no n=8 -v > gpurun_out/r2b_forced_no8.log 2>&1)
(timeout 600 python tests/model_checks.py --forced tdo -v > gpurun_out/r2b_forced_tdo_v.log 2>&1)
(timeout 600 python tests/model_checks.py --forced no n=4 > gpurun_out/r2b_forced_no4.log 2>&1)
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_dropin_scripts.py 2>&1 | tail -60 > gpurun_out/r2b_pytest.log
timeout 900 python -m pytest tests/test_dropin_scripts.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2b_dropin.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
grep -c FAIL gpurun_out/r2b_forced*.log; tail -5 gpurun_out/r2b_pytest.log; tail -3 gpurun_out/r2b_dropin.log
