#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x -k "varlen" -s > gpurun_out/o_tests_v.log 2>&1
echo "varlen tests rc=$?" >> gpurun_out/o_tests_v.log
grep -E "passed|failed|rc=|varlen|Error|error|assert" gpurun_out/o_tests_v.log | tail -30
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/o_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/o_tests.log
tail -5 gpurun_out/o_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err
tail -2 gpurun_out/o_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/o_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['settled']['value'],d['e2e']['value'],d['latency_batch8']['p50_ms'],d['omp_style_batch8']['value']);print(json.dumps(d['ragged_batch']))"
