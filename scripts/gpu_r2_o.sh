# round 2, call O (1 GPU): the sparse dynamics experiment -- new GPU tests and two 1M-node runs
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_experiment.py -q -x > gpurun_out/pytest_experiment.log 2>&1; echo "pytest experiment rc=$?"; tail -15 gpurun_out/pytest_experiment.log
timeout 300 python -m ndcn_b200.experiment --dynamics heat --network random --mean_degree 10 --n 1000000 --sampled_time equal \
  --niters 10 --test_freq 5 --method euler --hidden 20 > gpurun_out/r02_experiment_heat_er_1m.txt 2>&1; echo "heat 1M rc=$?"; grep -v Warn gpurun_out/r02_experiment_heat_er_1m.txt | tail -6
timeout 300 python -m ndcn_b200.experiment --dynamics gene --network power_law --n 1000000 --sampled_time irregular \
  --niters 10 --test_freq 5 --method euler --hidden 20 > gpurun_out/r02_experiment_gene_pl_1m.txt 2>&1; echo "gene 1M rc=$?"; grep -v Warn gpurun_out/r02_experiment_gene_pl_1m.txt | tail -6
