# multi-GPU run (under `gpurun --gpus N`): bash scripts/gpu_r2_multi.sh N "tag:ENV=V,... tag2:..." [steps]
# every entry = one torchrun of bench.py on N GPUs; the JSON lines land in gpurun_out/bench_r2_<N>gpu_<tag>.json
mkdir -p gpurun_out
np=${1:-2}
steps=${3:-10}
for entry in ${2:-default:}; do
  tag=${entry%%:*}
  envs=${entry#*:}
  envs=${envs//,/ }
  echo "== $np GPUs, $tag ($envs)"
  extra=""
  for kv in $envs; do case $kv in BENCH_EXTRA=*) extra="${kv#BENCH_EXTRA=}";; esac; done
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $np --steps $steps --warmup 3 $extra > gpurun_out/bench_r2_${np}gpu_$tag.json 2> gpurun_out/bench_r2_${np}gpu_$tag.err
  echo "rc=$?"
  python - <<PY
import json
try:
    txt = open('gpurun_out/bench_r2_${np}gpu_$tag.json').read()
    d = json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print("$tag", d["value"], "%.3f ms/step" % d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "parity", d.get("parity"))
    print({k: round(v / d["steps"], 3) for k, v in d["roofline"]["class_ms"].items()})
    print([ (r["rows"], r["stage"], r["gather"], r["exchange"]) for r in d.get("partition", {}).get("class_ms_per_step_by_rank", [])])
except Exception as e:
    print("$tag failed", e); print(open('gpurun_out/bench_r2_${np}gpu_$tag.err').read()[-3000:])
PY
done
