# round-1 evidence refresh: launch list, ncu --set full of the RHS kernels of one step, default bench
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01_umma.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_p.out 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_stage_ndcn_row|k_stage_gemm_umma" -s 14 -c 14 -o gpurun_out/prof_rhs_r01 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_p.out 2>&1; echo "full rc=$?"
timeout 900 python bench.py > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_p.json
