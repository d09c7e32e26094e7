mkdir -p gpurun_out
echo "== power law 100k (state 102 MB: L2 resident)"; timeout 300 python scripts/exp_kernels.py --nodes 100000 --spmm-only 2>&1 | grep spmm
echo "== power law 250k"; timeout 300 python scripts/exp_kernels.py --nodes 250000 --spmm-only 2>&1 | grep spmm
echo "== grid 1M (perfect locality)"; timeout 300 python scripts/exp_kernels.py --graph grid --spmm-only 2>&1 | grep spmm
echo "== ER 1M"; timeout 300 python scripts/exp_kernels.py --graph er --spmm-only 2>&1 | grep spmm
