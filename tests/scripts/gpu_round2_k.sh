#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_dropin_scripts.py ) 2>&1 | tail -40 > gpurun_out/r2k_pytest.log; tail -12 gpurun_out/r2k_pytest.log
