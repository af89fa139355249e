#!/bin/bash
# BASELINE.json configs[4] at full size: parity of sampled rows + a bench line (gliclass-qwen-1.5B architecture)
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
GLC_TEST_FULL=1 timeout 1500 python -m pytest tests/test_gpu_e2e.py -q -s -k "qwen2_1p5b" > gpurun_out/q_tests.log 2>&1
echo "rc=$?" >> gpurun_out/q_tests.log
timeout 900 python bench.py --arch qwen1.5b --batch 32 --seq 1024 --labels 20 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
grep -E "qwen|passed|failed|Error|rc=" gpurun_out/q_tests.log | tail; tail -c 3000 gpurun_out/q_bench.json; tail -3 gpurun_out/q_bench.err
