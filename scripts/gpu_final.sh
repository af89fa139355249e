# end-of-round evidence: full GPU suite, bench (both arms), ncu launch list, ncu --set full of attention / LN / GEMMs
mkdir -p gpurun_out
rm -f gpurun_out/f_*
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/f_tests.log
tail -n 3 gpurun_out/f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; tail -n 2 gpurun_out/f_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?" >> gpurun_out/f_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 25 -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_shift -s 4 -c 1 -o gpurun_out/f_attn python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_attn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:residual_ln_bulk -s 8 -c 1 -o gpurun_out/f_ln python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_ln.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_2cta -s 40 -c 4 -o gpurun_out/f_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_gemm.log 2>&1
tail -n 2 gpurun_out/f_bench.err; cat gpurun_out/f_bench.json gpurun_out/f_bench_ref.json
ls -la gpurun_out/ | grep " f_"
