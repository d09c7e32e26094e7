mkdir -p gpurun_out
for dbg in 0 64 68; do NDCN_UMMA_DBG=$dbg timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_r$dbg.json 2> gpurun_out/bench_r$dbg.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_r$dbg.json'))
print('dbg=$dbg', '%.3e'%d['value'], '%.2f ms/step'%d['ms_per_step'], d['roofline']['per_kernel']['gemm_epilogue'])
PY
done
timeout 900 python -m pytest tests/test_gpu_umma.py -x -q --timeout 900 -k full_size > gpurun_out/full_size.log 2>&1; echo "full-size test rc=$?"; tail -3 gpurun_out/full_size.log
