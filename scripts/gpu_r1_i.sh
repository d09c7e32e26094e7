mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/gpu_tests.log
timeout 900 python bench.py --steps 20 > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "bench rc=$?"; cat gpurun_out/bench_i.json
