# multi-GPU bench of the exchange schemes (run under `gpurun --gpus N`): bash scripts/gpu_multi.sh N "push feature halo" [extra bench flags]
mkdir -p gpurun_out
np=${1:-2}
for ex in ${2:-push}; do
tag=${ex}${4:-}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 10 --warmup 3 --exchange $ex --no-e2e $3 > gpurun_out/bench_${np}gpu_$tag.json 2> gpurun_out/bench_${np}gpu_$tag.err; echo "$np gpu $tag rc=$?"
python - <<PY
import json
try:
    txt=open('gpurun_out/bench_${np}gpu_$tag.json').read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print("$tag", "%.3e" % d["value"], "%.2f ms/step" % d["ms_per_step"], {k: round(v, 1) for k, v in d["roofline"]["class_ms"].items()}, d["solver"], d.get("partition", {}).get("row_bounds"))
except Exception as e:
    print("$tag failed", e); print(open('gpurun_out/bench_${np}gpu_$tag.err').read()[-2000:])
PY
done
