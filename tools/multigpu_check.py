"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multigpu_check.py

Residual samples are sharded by contiguous column blocks; every rank runs the
full mgvi_step on its shard.  Sums over residual samples (KL value, gradient input, mean metric
weight) are formed in a canonical order from all-gathered chunk partials (csrc/comm.cu), so the
sharded run must be BITWISE identical to the same step unsharded on one GPU: KL value, gradient,
mean metric product, updated centre, mnlp, Newton-CG trace and every residual sample.  Rank 0
also compares the residuals with the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = g.load_package()
    syn = pkg.synthetic
    ok = True
    # 1-D fields pair two residuals per complex transform: an even number per rank (16 / 8 ranks)
    counts = np.load(os.path.join(ROOT, "tests", "golden", "coal_mining_counts.npz"))["counts"]
    cases = [syn.field_config(dims, like, n, field_std=0.5)
             for dims, like, n in [((256, 256), syn.LIKE_NORMAL, 8), ((4096,), syn.LIKE_POISSON, 16),
                                   ((32, 32, 32), syn.LIKE_POISSON, 8), ((65536,), syn.LIKE_POISSON, 16)]]
    # the tutorial's hyperparameter field: per-sample metrics, canonical sum of n_theta-vectors across ranks
    cases.append(syn.hyper_field_config(counts, 8, start_scale=0.3))
    for cfg in cases:
        dims, n = cfg.get("dims", ("hyper", cfg["n_theta"])), cfg["n_residuals"]
        ctx = pkg.MGVIContext(device=local)
        box = [ctx.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.init_comm(rank, world, box[0])
        model = pkg.model_from_config(cfg, ctx)
        nl = n // world
        sl = slice(rank * nl, (rank + 1) * nl)
        Z1, Z2 = np.asfortranarray(cfg["Z1"][:, sl]), np.asfortranarray(cfg["Z2"][:, sl])
        conf = pkg.MGVIConfig(linsolver=pkg.B200CG(abstol=0.0, reltol=1e-12))
        res, center = pkg.mgvi_step(model, cfg["data"], nl, cfg["center"], conf, ctx, draws=(Z1, Z2))
        # sharded KL value/gradient and mean metric (all-reduced inside) must be identical on every rank
        R_loc = 0.5 * (res.samples[:, 0::2] - res.samples[:, 1::2])
        f, grad = pkg.kl_value_grad(model, cfg["center"], R_loc)
        u = np.random.default_rng(6).standard_normal(model.n_theta)
        mm = pkg.mean_metric_apply(model, cfg["center"], R_loc, u)
        t = torch.tensor([f, float(np.abs(grad).sum()), res.mnlp, float(np.abs(center).sum()), float(np.abs(mm).sum())],
                         dtype=torch.float64, device="cuda")
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        same = bool(torch.all(tmax == tmin))
        # gather the sharded samples in canonical column order on rank 0
        S_local = torch.from_numpy(np.ascontiguousarray(res.samples.T)).cuda()
        parts = [torch.empty_like(S_local) for _ in range(world)]
        dist.all_gather(parts, S_local)
        S = torch.cat(parts, dim=0).cpu().numpy().T
        if rank == 0:
            import oracle as O
            from helpers import TIGHT_O, oracle_model, rel
            ctx1 = pkg.MGVIContext(device=local)
            m1 = pkg.model_from_config(cfg, ctx1)
            res1, c1 = pkg.mgvi_step(m1, cfg["data"], n, cfg["center"], conf, ctx1, draws=(cfg["Z1"], cfg["Z2"]))
            Ro, _, _ = O.sample_residuals_cg(oracle_model(cfg), cfg["center"], cfg["Z1"], cfg["Z2"], **TIGHT_O)
            ref = dict(samples=O.build_samples(Ro, cfg["center"]))
            tr, tr1 = res.info["optimizer_output"], res1.info["optimizer_output"]
            R_all = 0.5 * (S[:, 0::2] - S[:, 1::2])
            f1, g1 = pkg.kl_value_grad(m1, cfg["center"], R_all)
            mm1 = pkg.mean_metric_apply(m1, cfg["center"], R_all, u)
            bit = dict(f=(f == f1), grad=bool(np.array_equal(grad, g1)), mean_metric=bool(np.array_equal(mm, mm1)),
                       center=bool(np.array_equal(center, c1)), samples=bool(np.array_equal(S, res1.samples)),
                       mnlp=(res.mnlp == res1.mnlp), trace=(tr == tr1))
            # residuals do not pass through the optimiser: tight against the oracle's CG solutions
            e_r = rel(R_all, 0.5 * (ref["samples"][:, 0::2] - ref["samples"][:, 1::2]))
            good = same and all(bit.values()) and e_r < 1e-10
            ok = ok and good
            print("dims %s world %d: ranks identical %s | sharded vs 1-GPU bitwise: %s | residuals vs oracle %.1e | "
                  "traces %s %s -> %s" % (dims, world, same, bit, e_r, tr["cg_updates"], tr1["cg_updates"],
                                          "PASS" if good else "FAIL"), flush=True)
            m1.close(); ctx1.close()
        model.close()
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if flag.item() != 1.0:
        sys.exit(1)
    if rank == 0:
        print("MULTIGPU CHECK PASS")


if __name__ == "__main__":
    main()
