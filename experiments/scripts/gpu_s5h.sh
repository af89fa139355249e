mkdir -p gpurun_out
rm -f gpurun_out/s5h_*
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "shift and not stream" > gpurun_out/s5h_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s5h_kernels.log
tail -n 5 gpurun_out/s5h_kernels.log
for h in 1 0; do
  echo "== SWAP=$h" >> gpurun_out/s5h_attn.log
  GLC_ATTN_SWAP=$h GLC_ATTN=shift timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s5h_attn.log 2>&1
done
GLC_ATTN=shift GLC_ATTN_TRACE=gpurun_out/s5h_trace.txt timeout 300 python scripts/bench_attn.py 64 512 12 1 >> gpurun_out/s5h_attn.log 2>&1
cat gpurun_out/s5h_attn.log; head -8 gpurun_out/s5h_trace.txt
