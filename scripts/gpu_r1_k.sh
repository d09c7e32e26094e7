mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -x -q --timeout 600 > gpurun_out/umma_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/umma_tests.log
for dbg in 0 4; do NDCN_UMMA_DBG=$dbg timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_k$dbg.json 2> gpurun_out/bench_k$dbg.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_k$dbg.json'))
print('dbg=$dbg', '%.3e'%d['value'], '%.2f ms/step'%d['ms_per_step'], d['roofline']['per_kernel'], 'e2e %.3e'%d['e2e']['value'], d['clocks'])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 24 --csv --log-file gpurun_out/launches_k.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_k.out 2>&1; echo "ncu list rc=$?"
