#!/bin/bash
# GPU-box visit 2: A/B of the latency experiments (persistent SM-affine patches, L1 prefetches, folded distant pow) in both sampler modes.
mkdir -p gpurun_out
timeout 300 python tools/shape_sweep.py --only old,fold,pers,pers_c4,pers_c7,pf1,pf3,pf4,pf7,pers_fold,pers_pf7_fold > gpurun_out/c2_sweep_fast.log 2>&1
cat gpurun_out/c2_sweep_fast.log
timeout 200 python tools/shape_sweep.py --flags 4 --only old,tex_p0i1,tex_p3i0,pers,pers_fold > gpurun_out/c2_sweep_tex.log 2>&1
cat gpurun_out/c2_sweep_tex.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:clouds_fast -s 2 -c 1 -o gpurun_out/c2_pers -f python tools/quick_perf.py --lib build/variants/lib_pers.so --only 1 > gpurun_out/c2_ncu_pers.log 2>&1
tail -2 gpurun_out/c2_ncu_pers.log
