#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/l_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/l_tests.log
python scripts/prof_batch.py base > gpurun_out/l_prof.txt 2>&1
bash scripts/gpu_profiles.sh > gpurun_out/l_profiles.log 2>&1
tail -4 gpurun_out/l_tests.log; cat gpurun_out/l_prof.txt; python -c "
import json;d=json.loads(open('gpurun_out/p_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['settled']['value'],d['e2e']['value'],d['latency_batch8']['p50_ms'],d['omp_style_batch8']['value'])"
ls -la gpurun_out | grep " p_"
