#!/bin/bash
# Validation call: everything the driver runs at round end, on the current code.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -25 > gpurun_out/r2h_pytest.log; tail -6 gpurun_out/r2h_pytest.log
( time timeout 600 python __graft_entry__.py smoke ) > gpurun_out/r2h_smoke.log 2>&1; tail -4 gpurun_out/r2h_smoke.log
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err; tail -3 gpurun_out/r2h_bench_ref.err
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -3 gpurun_out/r2h_bench.err; head -c 400 gpurun_out/r2h_bench.json
