#!/bin/bash
# GPU-box visit 3: new default (packed fp32 + folded distant pow) through the GPU suite; exact height band and direct-threshold A/B.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/c3_pytest.log 2>&1
tail -4 gpurun_out/c3_pytest.log
timeout 200 python tools/shape_sweep.py > gpurun_out/c3_sweep_fast.log 2>&1
cat gpurun_out/c3_sweep_fast.log
timeout 200 python tools/shape_sweep.py --flags 4 --only old,hb > gpurun_out/c3_sweep_tex.log 2>&1
cat gpurun_out/c3_sweep_tex.log
timeout 100 python tests/parity_report.py > gpurun_out/c3_parity_report.log 2>&1
cat gpurun_out/c3_parity_report.log | tail -20
