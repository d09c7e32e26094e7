cd tests/cuda && timeout 60 ./umma_probe; echo "probe rc=$?"; cd ../..
python -m pytest tests -m gpu -q --timeout 300 2>&1 > gpurun_out/gpu_tests_4.log; tail -6 gpurun_out/gpu_tests_4.log
