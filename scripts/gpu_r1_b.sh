mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -x -q --timeout 600 > gpurun_out/umma_tests.log 2>&1; echo "umma tests rc=$?"; tail -3 gpurun_out/umma_tests.log
timeout 600 python scripts/exp_kernels.py > gpurun_out/exp_kernels_b.log 2>&1; echo "exp rc=$?"; grep -v "^{" gpurun_out/exp_kernels_b.log | tail -12
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_b.json'))
print(d['value'], d['ms_per_step'], d['roofline']['class_ms'], d['roofline']['per_kernel'])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_gemm_umma -s 9 -c 1 -o gpurun_out/prof_umma_b -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_umma_b.out 2>&1; echo "ncu rc=$?"
