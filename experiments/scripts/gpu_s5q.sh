mkdir -p gpurun_out
rm -f gpurun_out/s5q_*
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 1200 python -m pytest tests/test_gpu_e2e.py -x -q > gpurun_out/s5q_e2e.log 2>&1; echo "rc=$?" >> gpurun_out/s5q_e2e.log
tail -n 3 gpurun_out/s5q_e2e.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s5q_bench.json 2> gpurun_out/s5q_bench.err; tail -3 gpurun_out/s5q_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s5q_bench.json"))
print("value", round(d['value'],1), "e2e", round(d['e2e']['value'],1), "sync", round(d['e2e']['synchronous_glc_run']['value'],1), d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'], d['latency_batch8'], d['omp_style_batch8']['value'])
PY
