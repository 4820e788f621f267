#!/bin/bash
# Round-end evidence on one B200: GPU suite, smoke, the bench lines (both arms), ncu launch list and full captures.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/f_pytest.log 2>&1
tail -3 gpurun_out/f_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; tail -1 gpurun_out/f_smoke.log
timeout 600 python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; tail -c 400 gpurun_out/f_bench_n1.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; tail -c 300 gpurun_out/f_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 3 --warmup 3 --no-extra --no-cpu > gpurun_out/f_launches_bench.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:clouds_fast_kernel -s 2 -c 1 -o gpurun_out/f_fast -f python tools/quick_perf.py --only 1 > gpurun_out/f_ncu_fast.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:clouds_fast_kernel -s 2 -c 1 -o gpurun_out/f_half -f python tools/quick_perf.py --only 1 --flags 8 > gpurun_out/f_ncu_half.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:clouds_fast_kernel -s 2 -c 1 -o gpurun_out/f_tex -f python tools/quick_perf.py --only 1 --flags 4 > gpurun_out/f_ncu_tex.log 2>&1
timeout 100 python tools/quick_perf.py > gpurun_out/f_quick_perf.log 2>&1; cat gpurun_out/f_quick_perf.log
timeout 200 python tests/parity_report.py > gpurun_out/f_parity_report.log 2>&1; tail -16 gpurun_out/f_parity_report.log
ls -la gpurun_out | head -40
