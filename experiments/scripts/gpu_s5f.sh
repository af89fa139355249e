# full GPU suite + bench + ncu evidence with the register-skew attention kernel as default (outputs under gpurun_out/q_*)
mkdir -p gpurun_out
rm -f gpurun_out/q_*
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/q_tests.log 2>&1; echo "rc=$?" >> gpurun_out/q_tests.log
tail -n 4 gpurun_out/q_tests.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?" >> gpurun_out/q_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/q_bench_ref.json 2> gpurun_out/q_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 13 -c 400 --csv --log-file gpurun_out/q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/q_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_shift -s 4 -c 1 -o gpurun_out/q_attn python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/q_ncu_attn.log 2>&1
GLC_ATTN=stream timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_stream -s 4 -c 1 -o gpurun_out/q_attn_stream python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/q_ncu_attn_stream.log 2>&1
tail -n 2 gpurun_out/q_bench.err; cat gpurun_out/q_bench.json gpurun_out/q_bench_ref.json
ls -la gpurun_out/ | grep " q_"
