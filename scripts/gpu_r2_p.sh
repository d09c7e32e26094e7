# round 2, call P (1 GPU): tensor-core weight-gradient reduction (mma.sync 3xTF32) vs the FP32-FMA tiles
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_surface.py -q -x -k "weight_grads or training or vjp" > gpurun_out/pytest_wg.log 2>&1; echo "pytest (mma) rc=$?"; tail -4 gpurun_out/pytest_wg.log
NDCN_WG_MMA=0 timeout 900 python -m pytest tests/test_gpu_surface.py -q -x -k "weight_grads" > gpurun_out/pytest_wg_fma.log 2>&1; echo "pytest (fma) rc=$?"; tail -2 gpurun_out/pytest_wg_fma.log
timeout 300 python scripts/exp_weight_grads.py > gpurun_out/r02_weight_grads_timing.txt 2>&1; NDCN_WG_MMA=0 timeout 300 python scripts/exp_weight_grads.py >> gpurun_out/r02_weight_grads_timing.txt 2>&1
cat gpurun_out/r02_weight_grads_timing.txt
timeout 400 python scripts/exp_training_step.py 2>/dev/null | grep "N=" > gpurun_out/r02_training_step_mma.txt; cat gpurun_out/r02_training_step_mma.txt
