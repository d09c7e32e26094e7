mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -x -q --timeout 900 > gpurun_out/umma_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/umma_tests.log
for dbg in 0 256; do NDCN_UMMA_DBG=$dbg timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_x$dbg.json 2> gpurun_out/bench_x$dbg.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_x$dbg.json'))
print('dbg=$dbg', '%.3e'%d['value'], '%.2f ms/step'%d['ms_per_step'], d['roofline']['per_kernel']['gemm_epilogue'], d['clocks']['sm_mhz'])
PY
done
for dbg in 0 256; do
NDCN_UMMA_DBG=$dbg timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 24 --csv --log-file gpurun_out/launches_x$dbg.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_x.out 2>&1; echo "dbg $dbg rc=$?"
done
timeout 900 ncu --set full --clock-control none -k regex:k_stage_gemm_umma -s 6 -c 1 -o gpurun_out/prof_err_r01 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_err.out 2>&1; echo "ncu err rc=$?"
