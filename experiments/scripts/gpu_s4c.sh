mkdir -p gpurun_out
rm -f gpurun_out/s4c_*
for f in 0 1 2 4 6 8 16 24 30 31; do
  echo "== flags $f" >> gpurun_out/s4c_flags.log
  GLC_ATTN_FLAGS=$f timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s4c_flags.log 2>&1
done
GLC_ATTN_TRACE=gpurun_out/s4c_trace.txt timeout 300 python scripts/bench_attn.py 64 512 12 1 >> gpurun_out/s4c_flags.log 2>&1
cat gpurun_out/s4c_flags.log | grep -v parity; cat gpurun_out/s4c_trace.txt
