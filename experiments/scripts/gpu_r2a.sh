#!/bin/bash
# round 2, GPU session A: new attention kernel parity + speed, full gpu suite, bench
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "attention" -x > gpurun_out/a_attn_tests.log 2>&1
echo "attn tests rc=$?" >> gpurun_out/a_attn_tests.log
timeout 300 python scripts/bench_attn.py 64 512 12 20 > gpurun_out/a_attn_bench.log 2>&1
timeout 300 python scripts/bench_attn.py 16 1024 12 20 >> gpurun_out/a_attn_bench.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/a_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/a_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?" >> gpurun_out/a_bench.err
tail -5 gpurun_out/a_attn_tests.log; cat gpurun_out/a_attn_bench.log; tail -15 gpurun_out/a_tests.log; tail -c 1500 gpurun_out/a_bench.json
