mkdir -p gpurun_out
rm -f gpurun_out/s5l_*
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "shift and not stream" > gpurun_out/s5l_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s5l_kernels.log
tail -n 5 gpurun_out/s5l_kernels.log
for h in 1 0 1 0; do
  echo "== POLY=$h" >> gpurun_out/s5l_attn.log
  GLC_ATTN_POLY=$h GLC_ATTN=shift timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s5l_attn.log 2>&1
done
GLC_ATTN=shift GLC_ATTN_TRACE=gpurun_out/s5l_trace.txt timeout 300 python scripts/bench_attn.py 64 512 12 1 >> gpurun_out/s5l_attn.log 2>&1
grep -v "mode" gpurun_out/s5l_attn.log; head -6 gpurun_out/s5l_trace.txt
