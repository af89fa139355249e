mkdir -p gpurun_out
rm -f gpurun_out/s5g_*
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "residual_ln or embed" > gpurun_out/s5g_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s5g_kernels.log
tail -n 5 gpurun_out/s5g_kernels.log
for b in 1 0; do
  echo "== BULK=$b" >> gpurun_out/s5g_ln.log
  GLC_LN_BULK=$b timeout 300 python scripts/bench_ln.py >> gpurun_out/s5g_ln.log 2>&1
done
cat gpurun_out/s5g_ln.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -x -q > gpurun_out/s5g_e2e.log 2>&1; echo "rc=$?" >> gpurun_out/s5g_e2e.log
tail -n 3 gpurun_out/s5g_e2e.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s5g_bench.json 2> gpurun_out/s5g_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s5g_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'], d['kernels']['residual_ln'], d['kernels']['attention'])
PY
