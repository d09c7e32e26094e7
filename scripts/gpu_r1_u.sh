mkdir -p gpurun_out
for np in 8 4; do for ex in feature halo; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 5 --warmup 3 --exchange $ex --no-e2e > gpurun_out/bench_${np}gpu_$ex.json 2> gpurun_out/bench_${np}gpu_$ex.err; echo "$np gpu $ex rc=$?"; python - <<PY
import json
txt=open('gpurun_out/bench_${np}gpu_$ex.json').read()
i=txt.find('{"metric"')
if i<0: print(txt[-1500:]); raise SystemExit
d=json.loads(txt[i:].splitlines()[0])
print('$np $ex', '%.3e'%d['value'], '%.2f ms/step'%d['ms_per_step'], d['roofline']['class_ms'], d['partition'])
PY
done; done
