#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "e4m3" -s > gpurun_out/n_tests_k.log 2>&1
echo "kernel tests rc=$?" >> gpurun_out/n_tests_k.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x -k "fp8 or unsupported" -s > gpurun_out/n_tests_e.log 2>&1
echo "e2e tests rc=$?" >> gpurun_out/n_tests_e.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --weights fp8 > gpurun_out/n_bench_fp8.json 2> gpurun_out/n_bench_fp8.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/n_bench_fp16.json 2> gpurun_out/n_bench_fp16.err
grep -E "passed|failed|rc=|e4m3|fp8-ffn|Error|error" gpurun_out/n_tests_k.log | tail -30; grep -E "passed|failed|rc=|fp8-ffn|Error" gpurun_out/n_tests_e.log | tail; for f in n_bench_fp8 n_bench_fp16; do tail -2 gpurun_out/$f.err; python -c "
import json;d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]);print('$f',d['value'],d['ms_per_step'],d['e2e']['value']);print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"; done
