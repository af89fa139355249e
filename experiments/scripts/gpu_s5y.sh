mkdir -p gpurun_out
rm -f gpurun_out/s5y_*
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 1200 python -m pytest tests/test_gpu_e2e.py -x -q -s > gpurun_out/s5y_e2e.log 2>&1; echo "rc=$?" >> gpurun_out/s5y_e2e.log
tail -n 3 gpurun_out/s5y_e2e.log
grep -h "max|d|" gpurun_out/s5y_e2e.log | sort -t= -k2 | tail -12
GLC_ATTN_G16=0 GLC_ATTN_C16=0 timeout 1200 python -m pytest tests/test_gpu_e2e.py -x -q -s -k "arch_parity or base_arch or large_arch" > gpurun_out/s5y_e2e32.log 2>&1
echo "--- fp32 C/G accumulators:"
grep -h "max|d|" gpurun_out/s5y_e2e32.log | tail -8
