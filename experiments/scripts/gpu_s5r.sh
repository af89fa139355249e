mkdir -p gpurun_out
rm -f gpurun_out/s5r_*
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or ln" > gpurun_out/s5r_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s5r_kernels.log
tail -n 5 gpurun_out/s5r_kernels.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -x -q > gpurun_out/s5r_e2e.log 2>&1; echo "rc=$?" >> gpurun_out/s5r_e2e.log
tail -n 3 gpurun_out/s5r_e2e.log
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s5r_$name.json 2> gpurun_out/s5r_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/s5r_$name.json"))
k=d['kernels']
print("$name", round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], "attn", k['attention']['ms_per_step'], "ln", k['residual_ln']['ms_per_step'], "out", k['gemm_out']['ms_per_step'], "ffn2", k['gemm_ffn2']['ms_per_step'])
PY
}
run fuse1 GLC_FUSE_RESID=1
run fuse0 GLC_FUSE_RESID=0
run fuse1b GLC_FUSE_RESID=1
run fuse0b GLC_FUSE_RESID=0
