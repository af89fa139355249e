# head variants + coalescing + async tests, then the bench
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 900 python -m pytest tests/test_gpu_e2e.py -x -q -k "head_variants or coalesced or submit_collect or golden_tiny or concurrent" > gpurun_out/r2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests.log
tail -n 25 gpurun_out/r2_tests.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?" >> gpurun_out/r2_bench.err
tail -n 3 gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench.json"))
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
print(d["latency_batch8"]); print(d["omp_style_batch8"])
for k,v in d["kernels"].items(): print("  ",k,v)
PY
