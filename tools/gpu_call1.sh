#!/bin/bash
# One GPU-box visit: GPU test suite, A/B of the staged kernel variants, the bench line, the ncu evidence.
# Everything lands in gpurun_out/ as it is produced, most important first.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_gpu.txt 2>&1
( time timeout 540 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.log 2>&1
tail -5 gpurun_out/c1_pytest.log
timeout 200 python tools/shape_sweep.py --only old,pk1,pk3,fold,fold_rec9,fold_rec7 > gpurun_out/c1_sweep_fast.log 2>&1
cat gpurun_out/c1_sweep_fast.log
timeout 100 python tools/shape_sweep.py --flags 4 --only old,texsmall > gpurun_out/c1_sweep_tex.log 2>&1
cat gpurun_out/c1_sweep_tex.log
timeout 300 python bench.py > gpurun_out/c1_bench_n1.json 2> gpurun_out/c1_bench_n1.err
cat gpurun_out/c1_bench_n1.json
timeout 60 python tools/noise_perf.py > gpurun_out/c1_noise_perf.log 2>&1
cat gpurun_out/c1_noise_perf.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:clouds_fast -s 2 -c 1 -o gpurun_out/c1_fast -f python tools/quick_perf.py --only 1 > gpurun_out/c1_ncu_full.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c1_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/c1_launches_bench.log 2>&1
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c1_bench_reference.json 2> gpurun_out/c1_bench_reference.err
tail -2 gpurun_out/c1_bench_reference.json
ls -la gpurun_out
