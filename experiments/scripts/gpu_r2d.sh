#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "attention and persist" -x > gpurun_out/d_attn_tests.log 2>&1
echo "attn tests rc=$?" >> gpurun_out/d_attn_tests.log
timeout 300 python scripts/bench_attn.py 64 512 12 20 > gpurun_out/d_attn_bench.log 2>&1
timeout 300 python scripts/bench_attn.py 16 1024 12 20 >> gpurun_out/d_attn_bench.log 2>&1
GLC_ATTN_MODE=0 GLC_ATTN=persist timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/d_attn_bench.log 2>&1
GLC_ATTN=persist timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
tail -25 gpurun_out/d_attn_tests.log; cat gpurun_out/d_attn_bench.log; python -c "
import json;d=json.loads(open('gpurun_out/d_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline_attention'],d['kernels'])"; tail -3 gpurun_out/d_bench.err
