mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -x -q --timeout 900 > gpurun_out/umma_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/umma_tests.log
for dbg in 0 128; do NDCN_UMMA_DBG=$dbg timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_s$dbg.json 2> gpurun_out/bench_s$dbg.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_s$dbg.json'))
print('dbg=$dbg', '%.3e'%d['value'], '%.2f ms/step'%d['ms_per_step'], d['roofline']['per_kernel']['gemm_epilogue'])
PY
done
