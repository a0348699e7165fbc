#!/bin/bash
# One gpurun call: GPU tests, bench, phase breakdown, ncu --set full of the hot kernels, ncu launch list of one bench step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
timeout ${T_TESTS:-900} python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests.log
timeout 420 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 240 python scripts/phase_times.py > gpurun_out/phase.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemv_kernel|gemm_tc|attn_fwd|decode_attn" -c 48 -f -o gpurun_out/prof_full python scripts/profile_kernels.py > gpurun_out/ncu_full.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --profile > gpurun_out/b_profile.log 2>&1
tail -3 gpurun_out/gpu_tests.log; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err; tail -4 gpurun_out/phase.log
