"""Generate the golden vectors under tests/golden/ from the CPU oracle.

The reference (Julia) cannot be executed in this environment, so these fixtures are outputs of
the ORACLE on small seeded inputs, not of MGVI.jl itself: they freeze the restatement (any later
change to oracle/ that moves a number shows up here) and give the GPU tests an input/output pair
that does not depend on importing the oracle.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import __graft_entry__ as g  # noqa: E402
import oracle as O  # noqa: E402
from helpers import TIGHT_O, oracle_model  # noqa: E402

syn = g.load_package().synthetic


def reference_fft_gp_config(n_residuals=4):
    """test/test_models/model_fft_gp.jl: 40 pixels, P = 50/(k^2.3 + 1), inv(DHT plan) = H/40,
    mean = exp(field), Normal(mean, 0.4).  The MersenneTwister draws of :75-76 are not
    reproducible without Julia; NumPy seeds 128 / 12 stand in."""
    n = 40
    k = np.array([i if i < n / 2 else n - i for i in range(n)], dtype=np.float64)
    A = 50.0 / (k ** 2.3 + 1.0)
    c = 1.0 / n
    xi_true = np.random.default_rng(128).standard_normal(n)
    X = np.fft.fft(A * xi_true)
    lam = np.exp(c * (X.real - X.imag))
    data = lam + 0.4 * np.random.default_rng(7).standard_normal(n)
    center = np.random.default_rng(12).standard_normal(n)
    rng = np.random.default_rng(42)
    return dict(kind="field", dims=(n,), amplitude=A, scale=c, offset=0.0, link=syn.LINK_EXP,
                likelihood=syn.LIKE_NORMAL, sigma=0.4, layout=syn.LAYOUT_INTERLEAVED, data=data, center=center,
                n_residuals=n_residuals, n_theta=n, n_lambda=2 * n,
                Z1=rng.standard_normal((n_residuals, 2 * n)).T, Z2=rng.standard_normal((n_residuals, n)).T)


CASES = {
    "fft_gp_40": reference_fft_gp_config,
    "field_1d_64_poisson": lambda: syn.field_config((64,), syn.LIKE_POISSON, 3, field_std=0.5),
    "field_2d_16x32_normal": lambda: syn.field_config((16, 32), syn.LIKE_NORMAL, 2, field_std=0.5,
                                                      layout=syn.LAYOUT_BLOCKED),
    "field_3d_8x8x16_poisson": lambda: syn.field_config((8, 8, 16), syn.LIKE_POISSON, 2, field_std=0.5),
    "polyfit_40": lambda: syn.polyfit_config(4, reference_grid=True),
}


def main():
    for name, mk in CASES.items():
        cfg = mk()
        om = oracle_model(cfg)
        c = cfg["center"]
        lin = om.linearize(c)
        V = np.random.default_rng(5).standard_normal((om.n_theta, 2))
        MV = np.stack([O.metric_apply(lin, V[:, i]) for i in range(2)], axis=1)
        R, iters, _ = O.sample_residuals_cg(om, c, cfg["Z1"], cfg["Z2"], **TIGHT_O)
        f = O.mean_neg_log_pstr(om, R, c)
        grad = O.kl_gradient(om, R, c)
        step = O.mgvi_step(om, c, Z1=cfg["Z1"], Z2=cfg["Z2"], **TIGHT_O)
        tr = step["trace"]
        arrays = {("cfg_" + k): np.asarray(v) for k, v in cfg.items()
                  if k not in ("kind", "x_groups", "xi_true", "true_params")}
        if cfg["kind"] == "poly":
            arrays["cfg_x_sizes"] = np.array([x.size for x in cfg["x_groups"]])
            arrays["cfg_x"] = np.concatenate(cfg["x_groups"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kind=cfg["kind"], V=V, MV=MV, R=R, iters=iters, f=f,
                            grad=grad, step_center=step["center"], step_samples=step["samples"],
                            step_mnlp=step["mnlp"], tr_cg_updates=np.array(tr.cg_updates),
                            tr_step_lengths=np.array(tr.step_lengths), tr_f_history=np.array(tr.f_history),
                            tr_f_calls=tr.f_calls, tr_g_calls=tr.g_calls, **arrays)
        print(name, "mnlp", step["mnlp"], "iters", iters)


if __name__ == "__main__":
    main()
