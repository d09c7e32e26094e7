# round 2, call A: full GPU test suite + bench (north star, configs 3/4) + slab probe rerun
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_r2a.json
NDCN_ERR_PREFIX=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e > gpurun_out/bench_r2a_noprefix.json 2> gpurun_out/bench_r2a_noprefix.err; echo "bench noprefix rc=$?"; python -c "
import json
for f in ['bench_r2a','bench_r2a_noprefix']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1]); print(f, d['ms_per_step'], d['roofline']['frac'], d['roofline']['class_ms'])
    except Exception as e: print(f,'failed',e)
"
timeout 300 python bench.py --config 3 --steps 100 --warmup 3 > gpurun_out/bench_r2a_cfg3.json 2> gpurun_out/bench_r2a_cfg3.err; echo "cfg3 rc=$?"; tail -c 1500 gpurun_out/bench_r2a_cfg3.json
python scripts/exp_slab_probe.py 1000000 power_law gpurun_out/graph.bin > /dev/null
./tests/cuda/slab_probe gpurun_out/graph.bin > gpurun_out/slab_probe_pl2.txt 2>&1; tail -12 gpurun_out/slab_probe_pl2.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
for mode in 1 14 17; do
  ncu --metrics $M --clock-control none -k regex:"k_slab|k_rowmajor" -s 3 -c 1 --csv --log-file gpurun_out/slab_ncu2_mode$mode.csv ./tests/cuda/slab_probe gpurun_out/graph.bin $mode > /dev/null 2>&1
done
rm -f gpurun_out/graph.bin
