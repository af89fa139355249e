#!/bin/bash
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 900 python bench.py --batch 64 --seq 1024 --labels 100 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(round(d['value'],1),round(d['e2e']['value'],1));print(json.dumps(d['ragged_batch']))"
