# round 2, call D: slab slice gather (in-process ranks), small solver after dyn1 change, full GPU suite
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
CUDA_MODULE_LOADING=EAGER timeout 900 python tests/push_inproc_worker.py > gpurun_out/push_inproc_slab.log 2>&1; echo "push inproc (slab) rc=$?"; grep CASE gpurun_out/push_inproc_slab.log | tail -20; tail -2 gpurun_out/push_inproc_slab.log
NDCN_FEAT_SLAB=0 CUDA_MODULE_LOADING=EAGER timeout 900 python tests/push_inproc_worker.py > gpurun_out/push_inproc_rows.log 2>&1; echo "push inproc (row layout) rc=$?"; tail -2 gpurun_out/push_inproc_rows.log
timeout 300 python scripts/exp_small_solver.py > gpurun_out/small_solver_timing.json 2> gpurun_out/small_solver_timing.err; echo "timing rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/small_solver_timing.json'))
for k,v in d.items(): print(k, v)
"
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -12 gpurun_out/pytest_gpu.log
