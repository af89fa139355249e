mkdir -p gpurun_out
rm -f gpurun_out/s5c_*
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "shift and stream" > gpurun_out/s5c_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s5c_kernels.log
tail -n 25 gpurun_out/s5c_kernels.log
GLC_ATTN=stream timeout 300 python scripts/bench_attn.py 64 512 12 20 > gpurun_out/s5c_attn.log 2>&1
GLC_ATTN=stream GLC_ATTN_TRACE=gpurun_out/s5c_trace.txt timeout 300 python scripts/bench_attn.py 64 512 12 1 >> gpurun_out/s5c_attn.log 2>&1
grep -v parity gpurun_out/s5c_attn.log; cat gpurun_out/s5c_trace.txt
