mkdir -p gpurun_out
rm -f gpurun_out/s6b_*
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "shift and xw" > gpurun_out/s6b_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s6b_kernels.log
tail -n 14 gpurun_out/s6b_kernels.log
for m in xw shift xw shift; do
  echo "== $m" >> gpurun_out/s6b_attn.log
  GLC_ATTN=$m timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s6b_attn.log 2>&1
done
grep -v "mode" gpurun_out/s6b_attn.log
