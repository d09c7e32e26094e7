# 1 GPU: full gpu suite, the in-process peer-push tests, and a short bench line (regression check of the store_y change)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --deselect tests/test_gpu_push.py > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/gpu_tests.log
timeout 600 python -m pytest tests/test_gpu_push.py -q --timeout 300 -x > gpurun_out/push_tests.log 2>&1; echo "push tests rc=$?"; tail -30 gpurun_out/push_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_push1.json 2> gpurun_out/bench_push1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_push1.json'))
print("%.3e" % d["value"], "%.2f ms/step" % d["ms_per_step"], d["roofline"]["per_kernel"], d["roofline"]["kernel"][:60])
PY
