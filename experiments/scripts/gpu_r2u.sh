#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "ln" > gpurun_out/u_tests.log 2>&1; tail -3 gpurun_out/u_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/u_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline_ln'],d['kernels']['residual_ln'])"
