mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -x -q --timeout 900 -k "external or rhs_vs_oracle" > gpurun_out/ext_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/ext_tests.log
for ex in feature halo; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --exchange $ex > gpurun_out/bench_2gpu_$ex.json 2> gpurun_out/bench_2gpu_$ex.err; echo "2gpu $ex rc=$?"; python - <<PY
import json
txt=open('gpurun_out/bench_2gpu_$ex.json').read()
i=txt.find('{"metric"')
if i<0: print(txt[-2000:]); raise SystemExit
d=json.loads(txt[i:].splitlines()[0])
print('$ex', '%.3e'%d['value'], '%.2f ms/step'%d['ms_per_step'], d['roofline']['class_ms'], d['partition'], d['solver'])
PY
tail -4 gpurun_out/bench_2gpu_$ex.err
done
