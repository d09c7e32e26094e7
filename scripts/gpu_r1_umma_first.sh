# first GPU pass of the chunk-major gather + tcgen05 stage kernel: parity, kernel timings, suite, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -x -q --timeout 600 > gpurun_out/umma_tests.log 2>&1; echo "umma tests rc=$?"; tail -15 gpurun_out/umma_tests.log
timeout 600 python scripts/exp_kernels.py > gpurun_out/exp_kernels.log 2>&1; echo "exp rc=$?"; tail -20 gpurun_out/exp_kernels.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -8 gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_umma.json 2> gpurun_out/bench_umma.err; echo "bench rc=$?"; cat gpurun_out/bench_umma.json; tail -3 gpurun_out/bench_umma.err
