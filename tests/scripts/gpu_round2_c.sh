#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=tests:rgb-proprioceptive-pose-estimator_b200:.
timeout 180 python tests/probe_cta_pair.py > gpurun_out/r2c_pair.log 2>&1; echo "probe rc $?" >> gpurun_out/r2c_pair.log
tail -4 gpurun_out/r2c_pair.log
if grep -q "FAILED: 0" gpurun_out/r2c_pair.log; then
  timeout 600 python tests/bench_layers.py 256 5001 0 5002 > gpurun_out/r2c_layers.log 2>&1
  tail -25 gpurun_out/r2c_layers.log
fi
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_dropin_scripts.py 2>&1 | tail -30 > gpurun_out/r2c_pytest.log
tail -4 gpurun_out/r2c_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c_smoke.log 2>&1; tail -3 gpurun_out/r2c_smoke.log
timeout 600 python bench.py --only-main --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; head -c 300 gpurun_out/r2c_bench.json
