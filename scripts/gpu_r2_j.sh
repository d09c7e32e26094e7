# round 2, call J (1 GPU): bench lines for profiles/, ncu evidence (CSV only: gpurun_out must stay < 64 MiB),
# row-chunked gather + stage kernel (Z kept in L2): correctness and timing; [N,1] dynamics with hub CTAs
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_solver.py tests/test_gpu_rhs.py tests/test_gpu_small.py -q > gpurun_out/pytest_sel.log 2>&1; echo "pytest sel rc=$?"; tail -4 gpurun_out/pytest_sel.log
NDCN_Z_CHUNK_ROWS=37888 timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_configs.py -q -k "full_size or config4 or umma_solver" > gpurun_out/pytest_zchunk.log 2>&1; echo "pytest zchunk rc=$?"; tail -4 gpurun_out/pytest_zchunk.log
for c in 0 18944 37888 75776; do
  NDCN_Z_CHUNK_ROWS=$c timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e > gpurun_out/bench_zchunk_$c.json 2> gpurun_out/bench_zchunk_$c.err; echo "chunk $c rc=$?"
done
python - <<'PY'
import json
for c in (0, 18944, 37888, 75776):
    try:
        d = json.loads([l for l in open('gpurun_out/bench_zchunk_%d.json' % c) if l.startswith('{')][-1])
        print('chunk', c, '%.3f ms/step' % d['ms_per_step'], {k: round(v / d['steps'], 3) for k, v in d['roofline']['class_ms'].items()}, d['clocks']['sm_mhz'], d['solver'])
    except Exception as e:
        print(c, 'FAILED', e)
PY
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
ncu --set full --clock-control none -k regex:"k_stage_ndcn_row|k_stage_gemm_umma" -s 14 -c 14 -o /tmp/r02_rhs_kernels python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r02_rhs_kernels.ncu-rep --page raw --csv > gpurun_out/r02_rhs_kernels_ncu_raw.csv 2>/dev/null; ls -la gpurun_out/r02_rhs_kernels_ncu_raw.csv
timeout 400 python bench.py --impl reference --ref-budget-s 60 > gpurun_out/r02_bench_reference_arm_60s.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref arm rc=$?"
du -sh gpurun_out
