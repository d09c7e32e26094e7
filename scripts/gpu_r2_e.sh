# round 2, call E (2 GPUs): the real multi-process IPC test + 2-GPU bench with the parity field
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_push.py -q -k two_gpus > gpurun_out/pytest_push_2gpu.log 2>&1; echo "2-GPU IPC test rc=$?"; tail -5 gpurun_out/pytest_push_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/push_worker.py > gpurun_out/push_worker_2gpu.log 2>&1; echo "push_worker rc=$?"; tail -12 gpurun_out/push_worker_2gpu.log
bash scripts/gpu_r2_multi.sh 2 "push: fpush_slab:BENCH_EXTRA=--exchange=fpush"
