# 1 GPU: in-process peer-push worker (push + feature push) and a short single-GPU bench (regression check)
mkdir -p gpurun_out
timeout 900 python tests/push_inproc_worker.py > gpurun_out/push_inproc.log 2>&1; echo "inproc rc=$?"; grep "CASE\|PUSH_INPROC" gpurun_out/push_inproc.log; grep -B2 -A12 "Traceback" gpurun_out/push_inproc.log | head -60
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_push1.json 2> gpurun_out/bench_push1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_push1.json'))
print("%.3e" % d["value"], "%.2f ms/step" % d["ms_per_step"], d["roofline"]["per_kernel"])
PY
