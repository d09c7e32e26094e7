# 2 GPUs: in-process push tests, the torchrun/IPC push test, and bench at 2 GPUs with push vs halo
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -8
timeout 600 python -m pytest tests/test_gpu_push.py -q --timeout 300 > gpurun_out/push_tests.log 2>&1; echo "push tests rc=$?"; tail -15 gpurun_out/push_tests.log
for ex in push halo; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --exchange $ex --no-e2e > gpurun_out/bench_2gpu_$ex.json 2> gpurun_out/bench_2gpu_$ex.err; echo "2 gpu $ex rc=$?"; tail -3 gpurun_out/bench_2gpu_$ex.err
done
python - <<'PY'
import json
for ex in ("push","halo"):
    try:
        d=json.load(open('gpurun_out/bench_2gpu_%s.json'%ex))
        print(ex, "%.3e" % d["value"], "%.2f ms/step" % d["ms_per_step"], d["roofline"]["class_ms"], d["solver"])
    except Exception as e: print(ex, "failed", e)
PY
