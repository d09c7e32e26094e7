mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -x -q --timeout 600 > gpurun_out/umma_tests.log 2>&1; echo "umma tests rc=$?"; tail -3 gpurun_out/umma_tests.log
NDCN_UMMA_DBG=4 timeout 600 python scripts/exp_kernels.py > gpurun_out/exp_kernels_e.log 2>&1; echo "exp rc=$?"; grep -v "^{" gpurun_out/exp_kernels_e.log | tail -12
