mkdir -p gpurun_out
rm -f gpurun_out/s5k_*
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "shift and not stream" > gpurun_out/s5k_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s5k_kernels.log
tail -n 5 gpurun_out/s5k_kernels.log
for h in 1 0 1 0; do
  echo "== OTMEM=$h" >> gpurun_out/s5k_attn.log
  GLC_ATTN_OTMEM=$h GLC_ATTN=shift timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s5k_attn.log 2>&1
done
echo "== OTMEM=1 S1024" >> gpurun_out/s5k_attn.log
GLC_ATTN_OTMEM=1 GLC_ATTN=shift timeout 300 python scripts/bench_attn.py 32 1024 12 10 >> gpurun_out/s5k_attn.log 2>&1
echo "== OTMEM=0 S1024" >> gpurun_out/s5k_attn.log
GLC_ATTN_OTMEM=0 GLC_ATTN=shift timeout 300 python scripts/bench_attn.py 32 1024 12 10 >> gpurun_out/s5k_attn.log 2>&1
GLC_ATTN=shift GLC_ATTN_TRACE=gpurun_out/s5k_trace.txt timeout 300 python scripts/bench_attn.py 64 512 12 1 >> gpurun_out/s5k_attn.log 2>&1
cat gpurun_out/s5k_attn.log; head -8 gpurun_out/s5k_trace.txt
