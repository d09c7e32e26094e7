mkdir -p gpurun_out
for dbg in 768 800; do
NDCN_UMMA_DBG=$dbg timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 12 --csv --log-file gpurun_out/launches_m$dbg.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_m.out 2>&1; echo "dbg $dbg rc=$?"
done
