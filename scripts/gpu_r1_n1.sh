mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_surface.py -x -q --timeout 600 > gpurun_out/surface_tests.log 2>&1; echo "surface rc=$?"; tail -25 gpurun_out/surface_tests.log
