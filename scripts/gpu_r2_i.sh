# round 2, call I (1 GPU): full GPU suite, final bench lines (north star, configs 4/5, ground-truth RHS), ncu evidence
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu_northstar.json 2> gpurun_out/r02_bench_1gpu_northstar.err; echo "bench rc=$?"
timeout 600 python bench.py --config 4 --no-cpu-baseline > gpurun_out/r02_bench_1gpu_config4.json 2> gpurun_out/r02_bench_1gpu_config4.err; echo "cfg4 rc=$?"
timeout 900 python bench.py --config 5 --steps 10 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_1gpu_config5.json 2> gpurun_out/r02_bench_1gpu_config5.err; echo "cfg5 rc=$?"
timeout 600 python bench.py --config 3 --steps 100 > gpurun_out/r02_bench_1gpu_config3.json 2> gpurun_out/r02_bench_1gpu_config3.err; echo "cfg3 rc=$?"
for r in heat gene mutual; do
  timeout 300 python bench.py --rhs $r --hidden 1 --dt 1e-4 --steps 50 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_1gpu_truth_$r.json 2> gpurun_out/r02_bench_1gpu_truth_$r.err; echo "$r rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_bench_1gpu_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('/')[-1], '%.3e' % d['value'], '%.3f ms/step' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], 'e2e', d['e2e'] and '%.3e' % d['e2e']['value'], d.get('gpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('value'), d['clocks'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_northstar.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_stage_ndcn_row|k_stage_gemm_umma" -s 14 -c 14 -o gpurun_out/r02_rhs_kernels python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
timeout 400 python bench.py --impl reference --ref-budget-s 60 > gpurun_out/r02_bench_reference_arm_60s.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref arm rc=$?"; tail -c 600 gpurun_out/r02_bench_reference_arm_60s.json
