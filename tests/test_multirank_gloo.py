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


# ------------------------------------------------------------------------------------------------
# The REAL library code (kernels + host orchestration, emulator build) sharded over `world` gloo
# ranks through mgvib200_comm_init_external: results must be BITWISE identical to the unsharded
# run -- sums over residual samples are formed in a canonical order from all-gathered chunk
# partials (csrc/comm.cu, engine.h: CanonShape), never inside the collective.
# ------------------------------------------------------------------------------------------------
def _gloo_allgather(world):
    import ctypes as C

    def allgather(send, recv, count):
        a = np.ctypeslib.as_array((C.c_double * count).from_address(send)).copy()
        t = torch.from_numpy(a)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        full = np.ascontiguousarray(np.concatenate([o.numpy() for o in outs]))
        C.memmove(recv, full.ctypes.data, 8 * count * world)
    return allgather


def _sharded_worker(rank, world, port, cases, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    pkg = g.load_package()
    syn = pkg.synthetic
    lib = pkg.Library(g.EMU_LIB)
    res = {}
    for dims, like, n in cases:
        cfg = syn.field_config(dims, getattr(syn, like), n, field_std=0.5)
        ctx = pkg.MGVIContext(library=lib)
        ctx.init_comm_external(rank, world, _gloo_allgather(world))
        model = pkg.model_from_config(cfg, ctx)
        nl = n // world
        sl = slice(rank * nl, (rank + 1) * nl)
        Z1, Z2 = np.asfortranarray(cfg["Z1"][:, sl]), np.asfortranarray(cfg["Z2"][:, sl])
        conf = pkg.MGVIConfig(linsolver=pkg.B200CG(abstol=0.0, reltol=1e-12), optimizer=pkg.NewtonCG(steps=2))
        r, center = pkg.mgvi_step(model, cfg["data"], nl, cfg["center"], conf, ctx, draws=(Z1, Z2))
        R_loc = 0.5 * (r.samples[:, 0::2] - r.samples[:, 1::2])
        f, grad = pkg.kl_value_grad(model, cfg["center"], R_loc)
        u = np.random.default_rng(6).standard_normal(model.n_theta)
        mm = pkg.mean_metric_apply(model, cfg["center"], R_loc, u)
        res[str(dims)] = dict(f=f, grad=grad, mm=mm, center=center, mnlp=r.mnlp, R=R_loc,
                              trace=r.info["optimizer_output"])
        model.close()
        ctx.close()
    out[rank] = res
    dist.destroy_process_group()


def _unsharded(cases):
    import __graft_entry__ as g
    pkg = g.load_package()
    syn = pkg.synthetic
    lib = pkg.Library(g.build_emu())
    ctx = pkg.MGVIContext(library=lib)
    res = {}
    for dims, like, n in cases:
        cfg = syn.field_config(dims, getattr(syn, like), n, field_std=0.5)
        model = pkg.model_from_config(cfg, ctx)
        conf = pkg.MGVIConfig(linsolver=pkg.B200CG(abstol=0.0, reltol=1e-12), optimizer=pkg.NewtonCG(steps=2))
        r, center = pkg.mgvi_step(model, cfg["data"], n, cfg["center"], conf, ctx, draws=(cfg["Z1"], cfg["Z2"]))
        R = 0.5 * (r.samples[:, 0::2] - r.samples[:, 1::2])
        f, grad = pkg.kl_value_grad(model, cfg["center"], R)
        u = np.random.default_rng(6).standard_normal(model.n_theta)
        mm = pkg.mean_metric_apply(model, cfg["center"], R, u)
        res[str(dims)] = dict(f=f, grad=grad, mm=mm, center=center, mnlp=r.mnlp, R=R, trace=r.info["optimizer_output"])
        model.close()
    ctx.close()
    return res


@pytest.mark.parametrize("world,cases", [
    (2, [((16, 16), "LIKE_POISSON", 4), ((64,), "LIKE_NORMAL", 4), ((8, 8, 8), "LIKE_POISSON", 2), ((6, 10), "LIKE_POISSON", 2)]),
    (4, [((8, 16), "LIKE_NORMAL", 8)]),
])
def test_sharded_library_is_bitwise_independent_of_rank_count(world, cases):
    ref = _unsharded(cases)
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29900 + (os.getpid() % 300) + world
    mp.spawn(_sharded_worker, args=(world, port, cases, out), nprocs=world, join=True)
    assert len(out) == world
    for dims, _like, n in cases:
        k = str(dims)
        nl = n // world
        for r in range(world):
            got = out[r][k]
            assert got["f"] == ref[k]["f"], (k, r)
            assert np.array_equal(got["grad"], ref[k]["grad"]), (k, r)
            assert np.array_equal(got["mm"], ref[k]["mm"]), (k, r)
            assert got["trace"] == ref[k]["trace"], (k, r, got["trace"], ref[k]["trace"])
            assert got["mnlp"] == ref[k]["mnlp"], (k, r)
            assert np.array_equal(got["center"], ref[k]["center"]), (k, r)
            assert np.array_equal(got["R"], ref[k]["R"][:, r * nl:(r + 1) * nl]), (k, r)


def test_unsupported_rank_counts_are_refused():
    """3 ranks do not divide the canonical chunking (gcd(n, 8) chunks): refused with a message, never
    silently order-dependent."""
    import __graft_entry__ as g
    pkg = g.load_package()
    lib = pkg.Library(g.build_emu())
    ctx = pkg.MGVIContext(library=lib)
    ctx.init_comm_external(0, 3, lambda s, r, c: None)
    syn = pkg.synthetic
    cfg = syn.field_config((8, 8), syn.LIKE_POISSON, 3, field_std=0.5)
    model = pkg.model_from_config(cfg, ctx)
    with pytest.raises(pkg.MGVIB200Error):
        # the (fake) all-gather returns garbage counts -> "same number of residual samples" or the chunk rule
        pkg.kl_value_grad(model, cfg["center"], 0.1 * cfg["Z2"][:, :1])
    with pytest.raises(pkg.MGVIB200Error):
        pkg.model_from_config(syn.polyfit_config(2, reference_grid=True), ctx)     # single-GPU model family
    model.close()
    ctx.close()
