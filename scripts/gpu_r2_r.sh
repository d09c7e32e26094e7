# round 2, call R (1 GPU): final check -- full GPU suite, smoke, default bench, reference arm (short budget), ncu of the dW kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu_northstar_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_bench_1gpu_northstar_final.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_baseline']['value'] if 'gpu_baseline' in d else None, d['cpu_baseline']['value'], d['clocks'])"
timeout 300 python bench.py --impl reference --ref-budget-s 20 --steps 2 --warmup 1 > gpurun_out/bench_ref_short.json 2> gpurun_out/bench_ref_short.err; echo "ref arm rc=$?"; tail -c 600 gpurun_out/bench_ref_short.json
timeout 300 ncu --set full --clock-control none -k regex:k_weight_grads_mma -c 1 -o /tmp/wg python scripts/exp_weight_grads.py > gpurun_out/ncu_wg.log 2>&1
ncu -i /tmp/wg.ncu-rep --page raw --csv > gpurun_out/r02_weight_grads_mma_ncu_raw.csv 2>/dev/null; wc -c gpurun_out/r02_weight_grads_mma_ncu_raw.csv
