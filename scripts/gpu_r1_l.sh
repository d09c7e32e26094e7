mkdir -p gpurun_out
for dbg in 0 256 512 768; do NDCN_UMMA_DBG=$dbg timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_l$dbg.json 2> gpurun_out/bench_l$dbg.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_l$dbg.json'))
print('dbg=$dbg', '%.3e'%d['value'], '%.2f ms/step'%d['ms_per_step'], d['roofline']['per_kernel']['gemm_epilogue'])
PY
done
