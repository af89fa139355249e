#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/j_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/j_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
GLC_NO_PDL=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/j_bench_nopdl.json 2> gpurun_out/j_bench_nopdl.err
tail -4 gpurun_out/j_tests.log; for f in j_bench j_bench_nopdl; do python -c "
import json;d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]);print('$f',d['value'],d['ms_per_step'],d['settled']['value'],d['e2e']['value'],d['latency_batch8']['p50_ms'],d['omp_style_batch8']['value'])"; done
