mkdir -p gpurun_out
rm -f gpurun_out/s4d_*
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "toeplitz" > gpurun_out/s4d_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s4d_kernels.log
tail -n 5 gpurun_out/s4d_kernels.log
if grep -q "rc=0" gpurun_out/s4d_kernels.log; then
  timeout 300 python scripts/bench_attn.py 64 512 12 20 > gpurun_out/s4d_attn.log 2>&1
  for f in 1 6 24 30 31; do
    echo "== flags $f" >> gpurun_out/s4d_attn.log
    GLC_ATTN_FLAGS=$f timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s4d_attn.log 2>&1
  done
  GLC_ATTN_TRACE=gpurun_out/s4d_trace.txt timeout 300 python scripts/bench_attn.py 64 512 12 1 >> gpurun_out/s4d_attn.log 2>&1
  grep -v parity gpurun_out/s4d_attn.log; cat gpurun_out/s4d_trace.txt
fi
