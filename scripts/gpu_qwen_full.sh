#!/bin/bash
# BASELINE.json configs[4] at full size (gliclass-qwen-1.5B architecture): parity of sampled rows (GLC_QWEN_PARITY=1) and
# bench lines in fp16 and with the opt-in e4m3 MLP
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
if [ "$GLC_QWEN_PARITY" = "1" ]; then
  GLC_TEST_FULL=1 timeout 1500 python -m pytest tests/test_gpu_e2e.py -q -s -k "qwen2_1p5b" > gpurun_out/q_tests.log 2>&1
  echo "rc=$?" >> gpurun_out/q_tests.log
  grep -E "qwen|passed|failed|Error|rc=" gpurun_out/q_tests.log | tail
fi
B="python bench.py --arch qwen1.5b --batch 32 --seq 1024 --labels 20 --steps 10 --warmup 3 --no-cpu-baseline --no-extras"
timeout 1500 $B > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
timeout 900 $B --weights fp8 > gpurun_out/q_bench_fp8.json 2> gpurun_out/q_bench_fp8.err
for f in q_bench q_bench_fp8; do tail -1 gpurun_out/$f.err; python -c "
import json;d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]);print('$f',round(d['value'],1),'texts/s',round(d['ms_per_step'],2),'ms e2e',round(d['e2e']['value'],1),'fwd frac',round(d['forward']['whole_forward_frac_of_tensor_peak'],3));print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})"; done
