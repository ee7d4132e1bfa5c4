"""Small run over every kernel family (2048-point columns with the tile-major scratch, 256-point
axes with 8-column tiles, the four-step 1-D path, the dense path) printing a few result norms.
Run it once on the GPU and once with MGVIB200_LIB=tests/simt_emu/libmgvib200_emu.so: the printed
numbers agree to the last digits (round 1: identical to 1e-16).  compute-sanitizer is closed on
the GPU pool, so this emulator cross-check plus the oracle parity tests stand in for memcheck."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

pkg = g.load_package()
syn = pkg.synthetic
ctx = pkg.MGVIContext()
for dims, like in [((2048, 32), syn.LIKE_NORMAL), ((32, 2048), syn.LIKE_POISSON), ((16, 256, 16), syn.LIKE_POISSON),
                   ((8192,), syn.LIKE_POISSON)]:
    cfg = syn.field_config(dims, like, 3, field_std=0.5)
    m = pkg.model_from_config(cfg, ctx)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(maxiters=3), ctx)
    V = np.random.default_rng(0).standard_normal((m.n_theta, 3))
    MV = s.metric_apply(V)
    R = pkg.sample_residuals(s, 3, draws=(cfg["Z1"], cfg["Z2"]))
    f, gr = pkg.kl_value_grad(m, cfg["center"], R)
    print(dims, float(np.abs(MV).max()), float(np.abs(R).max()), f, flush=True)
    m.close()
cfg = syn.dense_linear_config(200, 130, 2)
m = pkg.model_from_config(cfg, ctx)
s = pkg.ResidualSampler(m, cfg["center"], pkg.MatrixInversion(), ctx)
R = pkg.sample_residuals(s, 2, draws=cfg["Z"] if "Z" in cfg else None)
print("dense", float(np.abs(R).max()))
m.close()
print("SMALL HOT-PATH RUN DONE")
