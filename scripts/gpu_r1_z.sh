mkdir -p gpurun_out
timeout 300 python scripts/exp_umma_trace.py --mode store > gpurun_out/trace_store4.log 2>&1; tail -4 gpurun_out/trace_store4.log
timeout 900 python -m pytest tests/test_gpu_umma.py -x -q --timeout 900 > gpurun_out/umma_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/umma_tests.log
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_z.json'))
print('%.3e'%d['value'], '%.2f ms/step'%d['ms_per_step'], d['roofline']['per_kernel'], d['clocks']['sm_mhz'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 24 --csv --log-file gpurun_out/launches_z.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_z.out 2>&1; echo "list rc=$?"
