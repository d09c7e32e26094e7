# round 2, call C: small-solver tests + timing after the optimisation pass, previously failing tests, training paths
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_small.py -q > gpurun_out/pytest_small.log 2>&1; echo "pytest small rc=$?"; tail -15 gpurun_out/pytest_small.log
timeout 300 python scripts/exp_small_solver.py > gpurun_out/small_solver_timing.json 2> gpurun_out/small_solver_timing.err; echo "timing rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/small_solver_timing.json'))
for k,v in d.items(): print(k, v)
"; tail -5 gpurun_out/small_solver_timing.err
timeout 1800 python -m pytest tests/test_gpu_push.py tests/test_gpu_umma.py tests/test_gpu_surface.py tests/test_gpu_scripts.py tests/test_gpu_configs.py -q > gpurun_out/pytest_sel.log 2>&1; echo "pytest sel rc=$?"; tail -30 gpurun_out/pytest_sel.log
timeout 300 python scripts/exp_training_step.py > gpurun_out/training_step.txt 2>&1; tail -12 gpurun_out/training_step.txt
