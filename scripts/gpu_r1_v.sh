for h in 32 64 128; do echo "== H=$h"; timeout 300 python scripts/exp_kernels.py --hidden $h --spmm-only 2>&1 | grep "^spmm" | head -3; done
