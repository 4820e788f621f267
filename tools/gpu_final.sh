#!/bin/bash
# Round-end evidence on one B200: GPU suite, smoke, the bench lines (both arms, both samplers), ncu launch list and full captures.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_pytest.log 2>&1
tail -3 gpurun_out/f_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; tail -1 gpurun_out/f_smoke.log
timeout 300 python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; cat gpurun_out/f_bench_n1.json
timeout 300 python bench.py --sampler texture --steps 50 --warmup 5 > gpurun_out/f_bench_n1_tex.json 2> gpurun_out/f_bench_n1_tex.err; cat gpurun_out/f_bench_n1_tex.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; tail -1 gpurun_out/f_bench_reference.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/f_launches_bench.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:clouds_fast -s 2 -c 1 -o gpurun_out/f_fast -f python tools/quick_perf.py --only 1 > gpurun_out/f_ncu_fast.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:clouds_fast -s 2 -c 1 -o gpurun_out/f_tex -f python tools/quick_perf.py --only 1 --flags 4 > gpurun_out/f_ncu_tex.log 2>&1
timeout 100 python tools/quick_perf.py > gpurun_out/f_quick_perf.log 2>&1; cat gpurun_out/f_quick_perf.log
timeout 100 python tools/quick_perf.py --flags 2 --only 2 > gpurun_out/f_quick_perf_early.log 2>&1; cat gpurun_out/f_quick_perf_early.log
timeout 60 python tools/noise_perf.py > gpurun_out/f_noise_perf.log 2>&1
ls -la gpurun_out | head -40
