"""Load a golden fixture (tests/golden/*.npz) back into a config dict + expected outputs."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    cfg = {"kind": str(z["kind"])}
    gold = {}
    for k in z.files:
        if k.startswith("cfg_"):
            v = z[k]
            cfg[k[4:]] = v if v.ndim else v.item()
        elif k != "kind":
            gold[k] = z[k]
    if cfg["kind"] == "field":
        cfg["dims"] = tuple(int(d) for d in np.atleast_1d(cfg["dims"]))
    if cfg["kind"] == "poly":
        sizes = cfg.pop("x_sizes")
        x = cfg.pop("x")
        off = np.concatenate([[0], np.cumsum(sizes)])
        cfg["x_groups"] = [x[off[i]:off[i + 1]] for i in range(len(sizes))]
    for k in ("n_residuals", "n_theta", "n_lambda", "layout", "link", "likelihood"):
        if k in cfg:
            cfg[k] = int(cfg[k])
    return cfg, gold
