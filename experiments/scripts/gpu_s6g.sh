mkdir -p gpurun_out
rm -f gpurun_out/s6g_*
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" > gpurun_out/s6g_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s6g_kernels.log
tail -n 3 gpurun_out/s6g_kernels.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s6g_bench.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/s6g_bench.json')); k=d['kernels']; print(round(d['value'],1), d['ms_per_step'], d['clocks']['sm_mhz'], {n:k[n]['ms_per_step'] for n in ('gemm_qkv','gemm_out','gemm_ffn1','gemm_ffn2','attention','residual_ln')}, d['roofline']['frac'])"
