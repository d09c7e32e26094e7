mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_rhs.py -x -q --timeout 600 > gpurun_out/umma_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/umma_tests.log
NDCN_UMMA_DBG=4 timeout 600 python scripts/exp_kernels.py --spmm-only 2>&1 | grep "^spmm"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_stage_gather_v2|k_stage_ndcn_row" -c 6 -o gpurun_out/prof_gather_g -f python scripts/exp_kernels.py --quick --spmm-only > gpurun_out/ncu_gather_g.out 2>&1; echo "ncu rc=$?"
