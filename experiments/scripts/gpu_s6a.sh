mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
GLC_TIMING=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s6a_bench.json 2> gpurun_out/s6a_bench.err
grep "run_host B=64" gpurun_out/s6a_bench.err | tail -12
python -c "
import json; d=json.load(open('gpurun_out/s6a_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
