mkdir -p gpurun_out
rm -f gpurun_out/s5u_*
GLC_ATTN=shift GLC_ATTN_TRACE=gpurun_out/s5u_trace.txt timeout 300 python scripts/bench_attn.py 64 512 12 1 > gpurun_out/s5u_attn.log 2>&1
cat gpurun_out/s5u_trace.txt
