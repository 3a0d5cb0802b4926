#!/bin/bash
# Round 2, GPU call L (1 GPU): compute-sanitizer racecheck + memcheck on the round-2 kernels.
cd "$(dirname "$0")/.."
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all python scratch/gpu_race.py > gpurun_out/l_racecheck.log 2>&1
tail -5 gpurun_out/l_racecheck.log
timeout 1200 compute-sanitizer --tool memcheck python scratch/gpu_race.py > gpurun_out/l_memcheck.log 2>&1
tail -4 gpurun_out/l_memcheck.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -k "search_vs_oracle or nan_in_dem or plugin_template or err_mask or noise_level or spatial or serial" > gpurun_out/l_memcheck_tests.log 2>&1
tail -4 gpurun_out/l_memcheck_tests.log
