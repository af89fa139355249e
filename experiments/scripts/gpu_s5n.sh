mkdir -p gpurun_out
rm -f gpurun_out/s5n_*
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s5n_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s5n_tests.log
tail -n 4 gpurun_out/s5n_tests.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/s5n_bench.json 2> gpurun_out/s5n_bench.err; echo "bench rc=$?" >> gpurun_out/s5n_bench.err
GLC_ATTN=gather timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s5n_bench_gather.json 2> gpurun_out/s5n_bench_gather.err
python - <<'PY'
import json
for n in ["s5n_bench","s5n_bench_gather"]:
    d=json.load(open(f"gpurun_out/{n}.json"))
    print(n, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['kernels']['attention'], d['kernels']['residual_ln'], d['latency_batch8']['p50_ms'])
PY
