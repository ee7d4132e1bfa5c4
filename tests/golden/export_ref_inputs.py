"""Exports the INPUTS of the golden cases as raw Float64 files that oracle/make_ref_fixtures.jl (real
MGVI.jl v0.4.5, to be run by someone who has Julia) reads:  tests/golden/ref_inputs/<case>/meta.txt
(`name rows cols` per line) + <name>.f64 (column-major, little-endian).  Run:
    python tests/golden/export_ref_inputs.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from golden_io import load_golden  # noqa: E402

CASES = ("fft_gp_40", "polyfit_40")      # the two models the reference ships under test/test_models/


def write_case(path, arrays):
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "meta.txt"), "w") as f:
        for name, a in arrays.items():
            a = np.asarray(a, dtype="<f8")
            a = a.reshape(a.shape[0] if a.ndim else 1, -1)
            f.write("%s %d %d\n" % (name, a.shape[0], a.shape[1]))
            a.T.tofile(os.path.join(path, name + ".f64"))       # column-major on disk


def main():
    for name in CASES:
        cfg, gold = load_golden(name)
        write_case(os.path.join(HERE, "ref_inputs", name),
                   dict(center=cfg["center"], data=cfg["data"], Z1=cfg["Z1"], Z2=cfg["Z2"], V=gold["V"]))
        print(name, "->", os.path.join("tests", "golden", "ref_inputs", name))


if __name__ == "__main__":
    main()
