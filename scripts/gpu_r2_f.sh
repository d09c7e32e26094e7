# round 2, call F (8 GPUs): scaling lines with the parity field; slab vs row-layout slices at 8 and 4 GPUs
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
bash scripts/gpu_r2_multi.sh 8 "auto: fpush_rows:NDCN_FEAT_SLAB=0"
bash scripts/gpu_r2_multi.sh 4 "auto: fpush_rows:NDCN_FEAT_SLAB=0"
