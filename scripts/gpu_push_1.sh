# 1 GPU: full gpu suite incl. the in-process peer-push worker
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -25 gpurun_out/gpu_tests.log
