# quick loop: kernel parity + e2e parity (small) + bench
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/q_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/q_kernels.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -x -q -k "not base_arch" > gpurun_out/q_e2e.log 2>&1; echo "rc=$?" >> gpurun_out/q_e2e.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?" >> gpurun_out/q_bench.err
tail -n 4 gpurun_out/q_kernels.log gpurun_out/q_e2e.log gpurun_out/q_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/q_bench.json"))
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "gemm TF", round(d["roofline"]["achieved"],1), "frac_fwd", round(d["config"]["whole_forward_frac_of_tensor_peak"],3))
for k,v in d["kernels"].items(): print("  ",k,v)
print(d["clocks"])
PY
