#!/bin/bash
export GLC_MODEL_CACHE=/tmp/glc_models
python scripts/omp_probe.py 16 12 2>&1 | grep threads
python scripts/omp_probe.py 8 12 2>&1 | grep threads
python - <<'PY'
import json, subprocess
out = subprocess.run("python bench.py --steps 10 --warmup 3 --no-cpu-baseline", shell=True, capture_output=True, text=True).stdout
d = json.loads(out.strip().splitlines()[-1]); print('latency', d['latency_batch8']['p50_ms'], 'omp', d['omp_style_batch8']['value'], d['omp_style_batch8']['merged_launches'], d['omp_style_batch8']['requests_in_merged_launches'])
PY
