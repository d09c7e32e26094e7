mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -x -q --timeout 600 > gpurun_out/umma_tests.log 2>&1; echo "umma tests rc=$?"; tail -3 gpurun_out/umma_tests.log
timeout 300 python scripts/exp_umma_trace.py --mode store > gpurun_out/trace_store.log 2>&1; echo "trace rc=$?"; tail -12 gpurun_out/trace_store.log
timeout 300 python scripts/exp_umma_trace.py --mode stage4 > gpurun_out/trace_stage.log 2>&1; echo "trace rc=$?"; tail -12 gpurun_out/trace_stage.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c.json'))
print(d['value'], d['ms_per_step'], d['roofline']['class_ms'], d['roofline']['per_kernel'])
PY
