# final 1-GPU validation: the driver's own commands
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
