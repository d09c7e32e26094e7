# round 2, call H: TINY solver + persistent adjoint: tests, timing, training iteration
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_small.py -q -x > gpurun_out/pytest_small.log 2>&1; echo "pytest small rc=$?"; tail -15 gpurun_out/pytest_small.log
timeout 300 python scripts/exp_small_solver.py > gpurun_out/small_solver_timing.json 2> gpurun_out/small_solver_timing.err; echo "timing rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/small_solver_timing.json'))
for k,v in d.items(): print(k, v)
"
NDCN_TINY=0 timeout 300 python scripts/exp_small_solver.py > gpurun_out/small_solver_timing_notiny.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/small_solver_timing_notiny.json'))
for k,v in d.items():
    if 'persistent' in k: print('NDCN_TINY=0', k, v)
"
timeout 600 python scripts/exp_training_step.py > gpurun_out/training_step.txt 2>&1; grep "N=" gpurun_out/training_step.txt
timeout 900 python -m pytest tests/test_gpu_surface.py tests/test_gpu_solver.py tests/test_gpu_scripts.py -q > gpurun_out/pytest_sel.log 2>&1; echo "pytest sel rc=$?"; tail -8 gpurun_out/pytest_sel.log
