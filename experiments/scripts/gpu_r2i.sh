#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
NG=${NGPUS:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/i_bench_n$NG.json 2> gpurun_out/i_bench_n$NG.err
echo "rc=$?" >> gpurun_out/i_bench_n$NG.err
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -k "two_devices" > gpurun_out/i_tests_n$NG.log 2>&1
tail -3 gpurun_out/i_tests_n$NG.log
python -c "
import json;d=json.loads(open('gpurun_out/i_bench_n$NG.json').read().strip().splitlines()[-1]);print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['settled'],d.get('inprocess_sharded'))"; tail -5 gpurun_out/i_bench_n$NG.err
