# round 2, call N (8 GPUs): slab gather with owner-interleaved block order
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
bash scripts/gpu_r2_multi.sh 8 "auto_spread:"
bash scripts/gpu_r2_multi.sh 4 "auto_spread:"
