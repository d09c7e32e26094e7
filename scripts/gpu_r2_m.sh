# round 2, call M (1 GPU): owner-interleaved slab gather (in-process ranks), epilogue batch depth A/B, full GPU suite
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
CUDA_MODULE_LOADING=EAGER timeout 900 python tests/push_inproc_worker.py > gpurun_out/push_inproc_spread.log 2>&1; echo "push inproc (spread) rc=$?"; tail -2 gpurun_out/push_inproc_spread.log
for t in b2 b1; do
  dbg=0; [ $t = b1 ] && dbg=256
  NDCN_UMMA_DBG=$dbg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e > gpurun_out/bench_batch_$t.json 2> gpurun_out/bench_batch_$t.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_batch_$t.json') if l.startswith('{')][-1]); print('$t', d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'], {k: round(v/20,3) for k,v in d['roofline']['class_ms'].items()})"
done
timeout 300 python bench.py --rhs mutual --hidden 1 --dt 1e-4 --steps 50 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_1gpu_truth_mutual_final.json 2>/dev/null; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_bench_1gpu_truth_mutual_final.json') if l.startswith('{')][-1]); print('mutual', d['value'], d['ms_per_step'])"
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
