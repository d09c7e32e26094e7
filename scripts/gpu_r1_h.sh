mkdir -p gpurun_out
for v in 0 1 2 3 4 5; do echo "variant $v"; NDCN_ROW_VARIANT=$v timeout 300 python scripts/exp_kernels.py --spmm-only 2>&1 | grep "^spmm cw=-1"; done
echo "ER:"; for v in 0 1 3; do echo "variant $v"; NDCN_ROW_VARIANT=$v timeout 300 python scripts/exp_kernels.py --spmm-only --graph er 2>&1 | grep "^spmm cw=-1"; done
