mkdir -p gpurun_out
rm -f gpurun_out/s5j_*
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "residual_ln" > gpurun_out/s5j_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s5j_kernels.log
tail -n 5 gpurun_out/s5j_kernels.log
for b in 1 0; do
  echo "== BULK=$b" >> gpurun_out/s5j_ln.log
  GLC_LN_BULK=$b timeout 300 python scripts/bench_ln.py 131072 1024 >> gpurun_out/s5j_ln.log 2>&1
  GLC_LN_BULK=$b timeout 300 python scripts/bench_ln.py 32768 768 >> gpurun_out/s5j_ln.log 2>&1
done
cat gpurun_out/s5j_ln.log
