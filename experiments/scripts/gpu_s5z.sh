mkdir -p gpurun_out
rm -f gpurun_out/s5z_*
for m in 2048 4096 8192 16384 32768; do
  for b in 1 0; do
    echo "== M=$m BULK=$b" >> gpurun_out/s5z_ln.log
    GLC_LN_BULK=$b timeout 300 python scripts/bench_ln.py $m 768 >> gpurun_out/s5z_ln.log 2>&1
  done
done
cat gpurun_out/s5z_ln.log
