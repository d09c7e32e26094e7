# round 2, call B: GPU tests (new small-solver tests first), small-solver timing
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_small.py -x -q > gpurun_out/pytest_small.log 2>&1; echo "pytest small rc=$?"; tail -25 gpurun_out/pytest_small.log
timeout 300 python scripts/exp_small_solver.py > gpurun_out/small_solver_timing.json 2> gpurun_out/small_solver_timing.err; echo "timing rc=$?"; cat gpurun_out/small_solver_timing.json; tail -5 gpurun_out/small_solver_timing.err
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -40 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
