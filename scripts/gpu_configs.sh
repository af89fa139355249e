#!/bin/bash
# bench lines of the other BASELINE.json configs with the current build (C1 small, C3 large fp16 / e4m3-FFN, C4 reranker shape)
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
B="python bench.py --no-cpu-baseline --no-extras"
timeout 600 $B --arch small --batch 8 --seq 512 --labels 4 --steps 20 --warmup 5 > gpurun_out/c1.json 2> gpurun_out/c1.err
timeout 900 $B --arch large --batch 128 --seq 1024 --labels 50 --steps 5 --warmup 3 > gpurun_out/c3.json 2> gpurun_out/c3.err
timeout 900 $B --arch large --batch 128 --seq 1024 --labels 50 --steps 5 --warmup 3 --weights fp8 > gpurun_out/c3_fp8.json 2> gpurun_out/c3_fp8.err
timeout 900 $B --arch base --batch 512 --seq 1024 --labels 100 --steps 5 --warmup 3 > gpurun_out/c4.json 2> gpurun_out/c4.err
for f in c1 c3 c3_fp8 c4; do tail -1 gpurun_out/$f.err; python -c "
import json;d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]);print('$f',round(d['value'],1),'texts/s',round(d['ms_per_step'],2),'ms e2e',round(d['e2e']['value'],1),'fwd frac',round(d['forward']['whole_forward_frac_of_tensor_peak'],3),'gemm',round(d['roofline']['frac'],3),'attn',round(d['roofline_attention']['frac'],3))"; done
