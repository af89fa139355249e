mkdir -p gpurun_out
rm -f gpurun_out/s5t_*
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s5t_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s5t_tests.log
tail -n 3 gpurun_out/s5t_tests.log
grep -h "max|d|" gpurun_out/s5t_tests.log | head -5
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s5t_bench.json 2> gpurun_out/s5t_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s5t_bench.json"))
k=d['kernels']
print(round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e']['async_submit_collect']['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], "attn", k['attention']['ms_per_step'], "ln", k['residual_ln']['ms_per_step'])
PY
