# multi-GPU bench of both exchange schemes (run under `gpurun --gpus 8`): bash scripts/gpu_multi.sh "8 4 2"
mkdir -p gpurun_out
for np in ${1:-8}; do for ex in feature halo; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 5 --warmup 3 --exchange $ex --no-e2e > gpurun_out/bench_${np}gpu_$ex.json 2> gpurun_out/bench_${np}gpu_$ex.err; echo "$np gpu $ex rc=$?"
done; done
