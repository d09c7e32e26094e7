mkdir -p gpurun_out
python -m ndcn_b200._build > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_umma.py -x -q --timeout 600 > gpurun_out/umma_tests.log 2>&1; echo "umma tests rc=$?"; tail -4 gpurun_out/umma_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_gemm_umma -s 9 -c 2 -o gpurun_out/prof_umma_a -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_umma_a.out 2>&1; echo "ncu1 rc=$?"; tail -3 gpurun_out/ncu_umma_a.out
timeout 900 ncu --set full --clock-control none -k regex:"k_stage_ndcn_row|k_stage_gather_chunk" -c 8 -o gpurun_out/prof_gather_a -f python scripts/exp_kernels.py --quick > gpurun_out/ncu_gather_a.out 2>&1; echo "ncu2 rc=$?"; tail -3 gpurun_out/ncu_gather_a.out
ls -la gpurun_out
