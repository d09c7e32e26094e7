# round 2, call L (1 GPU): error-stage epilogue experiments (fast division build; the other epilogue batch depth), final test pass
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DNDCN_ERR_FAST_DIV -o ndcn_b200/libndcn_b200_fastdiv.so ndcn_b200/csrc/ndcn_api.cu
M=gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:"k_stage_gemm_umma" -s 14 -c 14 --csv --log-file gpurun_out/err_stage_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-e2e > /dev/null 2>&1
NDCN_B200_LIB=$PWD/ndcn_b200/libndcn_b200_fastdiv.so ncu --metrics $M --clock-control none -k regex:"k_stage_gemm_umma" -s 14 -c 14 --csv --log-file gpurun_out/err_stage_fastdiv.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-e2e > /dev/null 2>&1
NDCN_UMMA_DBG=256 ncu --metrics $M --clock-control none -k regex:"k_stage_gemm_umma" -s 14 -c 14 --csv --log-file gpurun_out/err_stage_otherbatch.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-e2e > /dev/null 2>&1
python - <<'PY'
import csv
for tag in ('default', 'fastdiv', 'otherbatch'):
    try:
        rows = [r for r in csv.reader(open('gpurun_out/err_stage_%s.csv' % tag)) if len(r) > 10]
        h = rows[0]; kn = h.index('Kernel Name'); v = h.index('Metric Value'); u = h.index('Metric Unit')
        print(tag, [(r[kn].split('<')[1].split('>')[0], r[v], r[u]) for r in rows[1:8]])
    except Exception as e:
        print(tag, 'failed', e)
PY
for t in default fastdiv; do
  lib=""; [ $t = fastdiv ] && lib=$PWD/ndcn_b200/libndcn_b200_fastdiv.so
  NDCN_B200_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e > gpurun_out/bench_errdiv_$t.json 2> gpurun_out/bench_errdiv_$t.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_errdiv_$t.json') if l.startswith('{')][-1]); print('$t', d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['solver'])"
done
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_small.py tests/test_gpu_solver.py -q > gpurun_out/pytest_sel.log 2>&1; echo "pytest sel rc=$?"; tail -4 gpurun_out/pytest_sel.log
for r in heat gene mutual; do
  timeout 300 python bench.py --rhs $r --hidden 1 --dt 1e-4 --steps 50 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_1gpu_truth_${r}_stream.json 2> gpurun_out/r02_bench_1gpu_truth_${r}_stream.err; echo "$r rc=$?"
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_bench_1gpu_truth_${r}_stream.json') if l.startswith('{')][-1]); print('$r', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['class_ms'])"
done
timeout 300 python scripts/exp_small_solver.py > gpurun_out/small_solver_timing.json 2> gpurun_out/small_solver_timing.err; python -c "
import json; d=json.load(open('gpurun_out/small_solver_timing.json'))
for k,v in d.items(): print(k, v)
"
timeout 600 python scripts/exp_training_step.py 2>/dev/null | grep "N=" | head -3
