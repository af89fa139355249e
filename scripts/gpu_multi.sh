#!/bin/bash
# torchrun bench on every GPU of the box (run through: scripts/gpurun_retry.sh <timeout> scripts/gpu_multi.sh <gpus>)
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/m_bench_n$N.json 2> gpurun_out/m_bench_n$N.err
echo "N=$N rc=$?"; tail -3 gpurun_out/m_bench_n$N.err; python -c "
import json;d=json.loads(open('gpurun_out/m_bench_n$N.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['settled']['value'],d['e2e']['value']);print(json.dumps(d['inprocess_sharded']))"
