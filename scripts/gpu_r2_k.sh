# round 2, call K (8 GPUs): BASELINE configs 4 (ER 1M) and 5 (power-law 4M) on 8 GPUs
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
bash scripts/gpu_r2_multi.sh 8 "config4:BENCH_EXTRA=--config=4" 10
bash scripts/gpu_r2_multi.sh 8 "config5:BENCH_EXTRA=--config=5" 5
