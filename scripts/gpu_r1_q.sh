mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_q.json'))
print('1M dopri5', '%.3e'%d['value'], '%.2f ms/step'%d['ms_per_step'], d['roofline']['per_kernel'])
PY
timeout 600 python bench.py --nodes 100489 --method rk4 --steps 100 --no-cpu-baseline --no-e2e > gpurun_out/bench_q_rk4.json 2> gpurun_out/bench_q_rk4.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_q_rk4.json'))
print('100k rk4', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], d['roofline']['per_kernel'])
PY
timeout 600 python bench.py --adaptive --steps 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_q_adapt.json 2> gpurun_out/bench_q_adapt.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_q_adapt.json'))
print('1M adaptive', '%.3e'%d['value'], '%.2f ms/attempt'%d['ms_per_step'], d['solver'])
PY
tail -3 gpurun_out/bench_q_adapt.err
