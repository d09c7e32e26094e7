# round 2: slab-gather probe (timing + ncu traffic), under gpurun on one B200
mkdir -p gpurun_out
python scripts/exp_slab_probe.py 1000000 power_law gpurun_out/graph.bin
./tests/cuda/slab_probe gpurun_out/graph.bin > gpurun_out/slab_probe_pl.txt 2>&1
cat gpurun_out/slab_probe_pl.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
for mode in ${MODES:-1 14 17 19}; do
  ncu --metrics $M --clock-control none -k regex:"k_slab|k_rowmajor" -s 3 -c 1 --csv --log-file gpurun_out/slab_ncu_mode$mode.csv ./tests/cuda/slab_probe gpurun_out/graph.bin $mode > /dev/null 2>&1
done
python scripts/exp_slab_probe.py 1000000 er gpurun_out/graph_er.bin
./tests/cuda/slab_probe gpurun_out/graph_er.bin > gpurun_out/slab_probe_er.txt 2>&1
cat gpurun_out/slab_probe_er.txt
rm -f gpurun_out/graph.bin gpurun_out/graph_er.bin
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
