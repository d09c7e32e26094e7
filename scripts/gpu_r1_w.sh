mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/gpu_tests.log
for v in 0 1 2 3; do echo "H=128 row variant $v"; NDCN_GATHER_CW=-1 NDCN_ROW128_VARIANT=$v timeout 300 python scripts/exp_kernels.py --hidden 128 --spmm-only 2>&1 | grep "^spmm cw=-1"; done
timeout 900 ncu --set full --clock-control none -k regex:"k_stage_gemm_umma<256, 2" -c 1 -o gpurun_out/prof_err_r01 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_err.out 2>&1; echo "ncu rc=$?"
