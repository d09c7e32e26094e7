mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_rhs.py -x -q --timeout 600 > gpurun_out/umma_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/umma_tests.log
for hub in 0 20000 40960 80000; do echo "hub rows $hub"; NDCN_HUB_ROWS=$hub timeout 300 python scripts/exp_kernels.py --spmm-only 2>&1 | grep "^spmm cw=-1"; done
for dbg in 0 16; do NDCN_UMMA_DBG=$dbg timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_j$dbg.json 2> gpurun_out/bench_j$dbg.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_j$dbg.json'))
print('dbg=$dbg', '%.3e'%d['value'], '%.2f ms/step'%d['ms_per_step'], d['roofline']['per_kernel'])
PY
done
NDCN_HUB_ROWS=40960 timeout 600 ncu --set full --clock-control none -k regex:"k_stage_ndcn_row" -c 2 -o gpurun_out/prof_gather_j -f python scripts/exp_kernels.py --quick --spmm-only > gpurun_out/ncu_gather_j.out 2>&1; echo "ncu rc=$?"
