"""Worker of tests/test_gpu_push.py::test_push_two_gpus_torchrun (launched by torch.distributed.run,
one process per GPU): the peer-push solve over CUDA IPC / NVLink against the single-GPU solve."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import ndcn_b200 as nb
    from ndcn_b200 import partition, solver
    from ndcn_b200 import workloads as wl

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    cases = [("push", 8192 * world + 5, 256, "dopri5"), ("push", 8192 * world + 5, 128, "rk4"), ("push", 3001, 20, "dopri5"),
             ("fpush", 9001, 256, "dopri5"), ("fpush", 5003, 128, "rk4")]
    for (scheme, n, H, method) in cases:
        phi = wl.graph_operator(wl.power_law_adjacency(n, 5, seed=1), "norm_lap")
        torch.manual_seed(0)
        lin = torch.nn.Linear(H, H)
        W, b = (lin.weight.detach() * 0.5).to(dev), lin.bias.detach().to(dev)
        x0 = torch.randn(n, H)
        if method == "dopri5":
            t = torch.tensor([0.0, 0.3, 1.0], dtype=torch.float64)
            kw = dict(method="dopri5", rtol=1e-2, atol=1e-3)
        else:
            t = torch.linspace(0, 1, 6, dtype=torch.float64)
            kw = dict(method="rk4")
        if scheme == "fpush":  # feature-sharded peer push: column slices instead of whole rows
            part = partition.FeaturePushPartition.build(phi, world, rank, dev, H, method)
        else:
            part = partition.PushPartition.build(phi, world, rank, dev, H, method)
        spec = nb.RhsSpec.ndcn(H, W, b)
        for rep in range(2):
            mine = nb.odeint_fused(part.graph, spec, x0[part.row0:part.row1].to(dev), t, peers=part, **kw)
        info = solver.last_solve_info
        objs = [None] * world  # ragged row blocks: gather through the host
        dist.all_gather_object(objs, mine.cpu())
        got = torch.cat(objs, dim=1)
        solver.release_workspaces()
        part.close()
        if rank == 0:
            g = nb.CsrGraph.from_scipy(phi, dev)
            ref = nb.odeint_fused(g, nb.RhsSpec.ndcn(H, W, b), x0.to(dev), t, **kw).cpu()
            ref_info = solver.last_solve_info
            err = float((got - ref).abs().max())
            close = torch.allclose(got, ref, rtol=1e-4, atol=1e-5)
            same = (info.nfe, info.n_accepted, info.n_rejected) == (ref_info.nfe, ref_info.n_accepted, ref_info.n_rejected)
            print("%s n=%d H=%d %s: max|diff|=%.3g close=%s counters %s vs %s" %
                  (scheme, n, H, method, err, close, (info.nfe, info.n_accepted, info.n_rejected),
                   (ref_info.nfe, ref_info.n_accepted, ref_info.n_rejected)), flush=True)
            ok = ok and close and same
        dist.barrier()
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, 0)
    if rank == 0 and float(flag.item()) == 1.0:
        print("PUSH_WORKER_OK", flush=True)
    dist.destroy_process_group()
    return 0 if float(flag.item()) == 1.0 else 1


if __name__ == "__main__":
    sys.exit(main())
