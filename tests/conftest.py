import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def emu_ctx(pkg):
    """Context on the g++/fiber-emulator build of the same sources (TEST INFRASTRUCTURE; the
    product never loads this library)."""
    import __graft_entry__ as g
    # MGVIB200_EMU_LIB: run the emulator tier against a variant build (tools/build_variants.py --emu)
    path = os.environ.get("MGVIB200_EMU_LIB") or g.build_emu()
    lib = pkg.Library(path)
    ctx = pkg.MGVIContext(library=lib)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def gpu_ctx(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ctx = pkg.MGVIContext(device=0)      # product library; raises if libmgvib200.so is missing
    yield ctx
    ctx.close()


def pytest_terminal_summary(terminalreporter):
    """Which regime every Newton-phase comparison ran in (helpers.PARITY_LOG): strict (free-running
    result at 1e-10 with equal traces) or teacher-forced (every Newton step from the oracle's x_n at
    1e-10, widened per quantity only to the oracle's own measured reproducibility)."""
    try:
        import helpers
    except Exception:       # noqa: BLE001
        return
    if not helpers.PARITY_LOG:
        return
    tr = terminalreporter
    tr.write_sep("-", "Newton-phase parity regimes (config -> regime, worst relative error)")
    for label, regime, worst in helpers.PARITY_LOG:
        tr.write_line("  %-30s %s" % (label, regime if regime != "strict" else "strict (free-running end to end) %.1e" % worst))
