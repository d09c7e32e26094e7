"""Golden stdout of the UNMODIFIED reference scripts on the reference's own CPU path (build container only).

    python tests/golden/make_script_golden.py

Runs ``heat_dynamics.py`` (BASELINE config 1, 40 iterations) and ``dgnn.py`` (BASELINE config 2 = README flags,
3 epochs) from /root/reference through ``oracle/script_runner.py`` (torch/numpy seeded with 0; scripts untouched)
and stores the numbers they print in ``tests/golden/script_runs.json``.  ``tests/test_gpu_scripts.py`` runs the same
files from ``baseline/_ref`` through ``python -m ndcn_b200.run`` on the B200 and compares.
"""
from __future__ import annotations

import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.environ.get("NDCN_REFERENCE_ROOT", "/root/reference")

HEAT_ARGS = ["--network", "grid", "--T", "5", "--sampled_time", "equal", "--baseline", "ndcn", "--niters", "40"]
GENE_ARGS = ["--network", "grid", "--T", "5", "--sampled_time", "equal", "--baseline", "ndcn", "--niters", "20",
             "--method", "rk4", "--sparse"]
DGNN_ARGS = ["--dataset", "cora", "--model", "differential_gcn", "--iter", "1", "--dropout", "0", "--hidden", "256",
             "--T", "1.2", "--time_tick", "16", "--epochs", "3", "--weight_decay", "0.024", "--no_control",
             "--method", "dopri5", "--alpha", "0", "--seed", "0"]

ITER_RE = re.compile(r"Iter (\d+)\| Train Loss ([\d.]+)\(([\d.]+) Relative\) \| Test Loss ([\d.]+)\(([\d.]+) Relative\)")
EPOCH_RE = re.compile(r"Epoch: (\d+) loss_train: ([\d.]+) acc_train: ([\d.]+) loss_val: ([\d.]+) acc_val: ([\d.]+)")
TEST_RE = re.compile(r"Test set results: loss= ([\d.]+) accuracy= ([\d.]+)")


def parse_dynamics(text):
    out = []
    for m in ITER_RE.finditer(text):
        out.append({"iter": int(m.group(1)), "train": float(m.group(2)), "train_rel": float(m.group(3)),
                    "test": float(m.group(4)), "test_rel": float(m.group(5))})
    return out


def parse_dgnn(text):
    ep = [{"epoch": int(m.group(1)), "loss_train": float(m.group(2)), "acc_train": float(m.group(3)),
           "loss_val": float(m.group(4)), "acc_val": float(m.group(5))} for m in EPOCH_RE.finditer(text)]
    t = TEST_RE.search(text)
    return {"epochs": ep, "test_loss": float(t.group(1)), "test_acc": float(t.group(2))}


def run_ref(script, args):
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "script_runner.py"), REF, "0", "-", script] + args
    res = subprocess.run(cmd, capture_output=True, text=True, check=True)
    return res.stdout


def main():
    out = {"seed": 0, "note": "reference scripts unmodified, torch %s CPU" % __import__("torch").__version__}
    out["heat"] = {"args": HEAT_ARGS, "cpu_flag": ["--gpu", "-1"], "lines": parse_dynamics(run_ref("heat_dynamics.py", HEAT_ARGS + ["--gpu", "-1"]))}
    out["gene_rk4_sparse"] = {"args": GENE_ARGS, "cpu_flag": ["--gpu", "-1"],
                              "lines": parse_dynamics(run_ref("gene_dynamics.py", GENE_ARGS + ["--gpu", "-1"]))}
    out["dgnn"] = {"args": DGNN_ARGS, "cpu_flag": ["--no-cuda"], **parse_dgnn(run_ref("dgnn.py", DGNN_ARGS + ["--no-cuda"]))}
    with open(os.path.join(HERE, "script_runs.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
