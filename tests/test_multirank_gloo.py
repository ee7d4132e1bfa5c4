"""world_size-2 `gloo` test of the sample-sharded Newton phase, on CPU.

What the N > 1 path relies on, beyond the single-GPU kernels, is algebra: the KL value, the
gradient input  sum_p g'(s_p) dl/dlambda(p)  and the mean metric weight  w_bar  are SUMS over
residual samples, so per-rank partial sums over contiguous column shards, all-reduced, reproduce
the unsharded quantities, and every rank then runs the identical scalar recurrences.  This test
does exactly that with the oracle's pieces as the per-rank "kernels" and torch.distributed/gloo
as the collective; the NCCL twin of it runs on GPUs in tools/multigpu_check.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    import oracle as O
    from helpers import oracle_model
    pkg = g.load_package()
    syn = pkg.synthetic
    cfg = syn.field_config((16, 16), syn.LIKE_POISSON, 6, field_std=0.5)
    om = oracle_model(cfg)
    c = cfg["center"]
    rng = np.random.default_rng(3)
    R = 0.1 * rng.standard_normal((om.n_theta, 6))
    n = R.shape[1]
    nl = n // world
    mine = range(rank * nl, (rank + 1) * nl)            # contiguous column shard, as bench.py does
    # per-rank partial sums over the local '+/-' points
    f_part = 0.0
    q_part = np.zeros(om.N)
    w_part = np.zeros(om.N)
    for i in mine:
        for sgn in (1.0, -1.0):
            p = c + sgn * R[:, i]
            lam = om.rate(p)
            f_part += -(O.posterior_loglike(om, p))
            q_part += cfg["data"] - lam                  # Poisson + exp link: g' dl/dlambda = d - lambda
        w_part += om.rate(c + R[:, i])                   # Poisson + exp link: w = lambda
    buf = torch.from_numpy(np.concatenate([q_part, w_part, [f_part, float(len(mine))]]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    tot = buf.numpy()
    n_glob = int(round(tot[-1]))
    f = tot[-2] / (2 * n_glob)
    grad = c - om.c * (om.A * om._H(tot[: om.N])) / (2 * n_glob)
    wbar = tot[om.N: 2 * om.N] / n_glob
    u = np.random.default_rng(4).standard_normal(om.N)
    mm = u + om.c ** 2 * (om.A * om._H(wbar * om._H(om.A * u)))
    # unsharded references
    f0 = O.mean_neg_log_pstr(om, R, c)
    g0 = O.kl_gradient(om, R, c)
    mm0 = O.mean_metric_apply(om, R, c, u)
    res = dict(n=n_glob, f=abs(f - f0) / abs(f0), g=np.linalg.norm(grad - g0) / np.linalg.norm(g0),
               mm=np.linalg.norm(mm - mm0) / np.linalg.norm(mm0), f_val=f)
    out[rank] = res
    dist.destroy_process_group()


def test_sharded_newton_phase_reduces_to_unsharded():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        assert out[r]["n"] == 6
        assert out[r]["f"] < 1e-12 and out[r]["g"] < 1e-12 and out[r]["mm"] < 1e-12, dict(out[r])
    assert out[0]["f_val"] == out[1]["f_val"]            # bitwise identical reduced value on every rank


def test_shard_layout_matches_bench(pkg):
    # rank g of G owns residuals [g n/G, (g+1) n/G): every residual exactly once, in order
    n, G = 32, 8
    owned = [i for g_ in range(G) for i in range(g_ * (n // G), (g_ + 1) * (n // G))]
    assert owned == list(range(n))
