# session-4 re-validation: full GPU parity suite, smoke, bench, attention what-if flags
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
nproc > gpurun_out/s4a_host.txt; nvidia-smi -L >> gpurun_out/s4a_host.txt
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/s4a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4a_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s4a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s4a_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s4a_bench.json 2> gpurun_out/s4a_bench.err; echo "bench rc=$?" >> gpurun_out/s4a_bench.err
for f in 0 1 2 4 6 7 8 15 16; do
  echo "== flags $f" >> gpurun_out/s4a_attn_flags.log
  GLC_ATTN_FLAGS=$f timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s4a_attn_flags.log 2>&1
done
echo "== no flags" >> gpurun_out/s4a_attn_flags.log
timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s4a_attn_flags.log 2>&1
GLC_ATTN_TRACE=gpurun_out/s4a_attn_trace.txt timeout 300 python scripts/bench_attn.py 64 512 12 1 >> gpurun_out/s4a_attn_flags.log 2>&1
tail -n 5 gpurun_out/s4a_pytest.log gpurun_out/s4a_smoke.log gpurun_out/s4a_bench.err; cat gpurun_out/s4a_bench.json; cat gpurun_out/s4a_attn_flags.log
