#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "attention" -x > gpurun_out/c_attn_tests.log 2>&1
echo "attn tests rc=$?" >> gpurun_out/c_attn_tests.log
timeout 300 python scripts/bench_attn.py 64 512 12 20 > gpurun_out/c_attn_bench.log 2>&1
timeout 300 python scripts/bench_attn.py 16 1024 12 20 >> gpurun_out/c_attn_bench.log 2>&1
GLC_ATTN=rows GLC_ATTN_TRACE=gpurun_out/c_trace_s512.txt timeout 300 python scripts/bench_attn.py 64 512 12 1 > gpurun_out/c_trace.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
tail -4 gpurun_out/c_attn_tests.log; cat gpurun_out/c_attn_bench.log; cat gpurun_out/c_trace_s512.txt | head -60; python -c "
import json;d=json.loads(open('gpurun_out/c_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['settled'],d['roofline_attention'],d['kernels'])"
