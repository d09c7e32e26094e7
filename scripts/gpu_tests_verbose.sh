python -m pytest tests -m gpu -q --timeout 300 2>&1 > gpurun_out/gpu_tests_3.log
tail -5 gpurun_out/gpu_tests_3.log
timeout 900 python bench.py --steps 10 > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err
cat gpurun_out/bench_1m.json; tail -3 gpurun_out/bench_1m.err
timeout 900 python bench.py --steps 10 --layout degree --no-cpu-baseline --no-e2e > gpurun_out/bench_1m_degree.json 2> gpurun_out/bench_1m_degree.err
cat gpurun_out/bench_1m_degree.json
timeout 900 python bench.py --steps 10 --graph er --no-cpu-baseline --no-e2e > gpurun_out/bench_1m_er.json 2> gpurun_out/bench_1m_er.err
cat gpurun_out/bench_1m_er.json
