# bench + reference arm + ncu evidence for the current build (outputs under gpurun_out/p_*)
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; echo "bench rc=$?" >> gpurun_out/p_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/p_bench_ref.json 2> gpurun_out/p_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 13 -c 400 --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/p_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_2cta -s 40 -c 4 -o gpurun_out/p_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/p_ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fused -s 4 -c 1 -o gpurun_out/p_attn python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/p_ncu_attn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:residual_ln -s 8 -c 1 -o gpurun_out/p_ln python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/p_ncu_ln.log 2>&1
tail -n 2 gpurun_out/p_bench.err; cat gpurun_out/p_bench.json gpurun_out/p_bench_ref.json
ls -la gpurun_out/ | grep " p_"
