#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 300 python scripts/bench_attn.py 64 512 12 20 > gpurun_out/f_attn_bench.log 2>&1
timeout 300 python scripts/bench_attn.py 16 1024 12 20 >> gpurun_out/f_attn_bench.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/f_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/f_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
cat gpurun_out/f_attn_bench.log; tail -6 gpurun_out/f_tests.log; python -c "
import json;d=json.loads(open('gpurun_out/f_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['settled'],d['roofline_attention']['us_per_launch'],d['latency_batch8'],d['inprocess_sharded'],d['kernels'])"; tail -3 gpurun_out/f_bench.err
