# round 2, call G (8 GPUs): slab gather with the per-rank start rotation
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
bash scripts/gpu_r2_multi.sh 8 "auto_rot:"
bash scripts/gpu_r2_multi.sh 4 "auto_rot:"
