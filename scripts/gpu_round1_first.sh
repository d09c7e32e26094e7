set -x
python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/gpu_tests_2.log
timeout 600 python bench.py --nodes 100000 --steps 5 --no-cpu-baseline > gpurun_out/bench_100k.json 2> gpurun_out/bench_100k.err
timeout 900 python bench.py --steps 10 > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_ndcn_gemm -s 9 -c 2 -o gpurun_out/prof_stage_r01 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.out 2>&1
tail -5 gpurun_out/gpu_tests_2.log; cat gpurun_out/bench_100k.json; cat gpurun_out/bench_1m.json; tail -3 gpurun_out/bench_1m.err
