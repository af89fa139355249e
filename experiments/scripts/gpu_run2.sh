# parity suite the way the driver runs it, then bench + profiles
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?" >> gpurun_out/r2_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 13 -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16 -s 40 -c 4 -o gpurun_out/r2_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fused -s 4 -c 1 -o gpurun_out/r2_attn python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_attn.log 2>&1
tail -n 5 gpurun_out/r2_pytest.log gpurun_out/r2_smoke.log gpurun_out/r2_bench.err; cat gpurun_out/r2_bench.json gpurun_out/r2_bench_ref.json
ls -la gpurun_out/
