mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/gpu_tests.log
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_y.json'))
print('%.3e'%d['value'], '%.2f ms/step'%d['ms_per_step'], d['roofline']['per_kernel'], d['clocks']['sm_mhz'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 24 --csv --log-file gpurun_out/launches_y.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_y.out 2>&1; echo "list rc=$?"
