"""The product library loads on a box without a GPU and exports every symbol the header
declares; the host API rejects what has no device path.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mgvi_b200.h")).read()
    return sorted(set(re.findall(r"\b(mgvib200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    import __graft_entry__ as g
    path = g.build_cuda()
    dll = ctypes.CDLL(path)
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(dll, n), "missing export %s" % n
    assert set(names) == set(pkg.capi.EXPORTS), set(names) ^ set(pkg.capi.EXPORTS)
    dll.mgvib200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in dll.mgvib200_version()


def test_cuda_objects_are_sm100a():
    import subprocess
    import __graft_entry__ as g
    path = g.build_cuda()
    out = subprocess.run(["cuobjdump", "--list-elf", path], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out, out


def test_missing_library_fails_loudly(pkg, tmp_path):
    with pytest.raises(pkg.MGVIB200Error, match="no CPU fallback"):
        pkg.Library(str(tmp_path / "nope.so"))


def test_emu_library_exports_same_abi(pkg):
    import __graft_entry__ as g
    dll = ctypes.CDLL(g.build_emu())
    for n in _declared_symbols():
        assert hasattr(dll, n), n


def test_host_api_rejections(pkg, emu_ctx):
    syn = pkg.synthetic
    with pytest.raises(pkg.MGVIB200Error, match="declarative"):
        pkg.ResidualSampler(lambda p: p, np.zeros(3), pkg.B200CG(), emu_ctx)
    with pytest.raises(pkg.MGVIB200Error, match="declarative"):
        pkg.mgvi_step(lambda p: p, None, 2, np.zeros(3), pkg.MGVIConfig(), emu_ctx)
    cfg = syn.field_config((16,), syn.LIKE_POISSON, 2)
    m = pkg.model_from_config(cfg, emu_ctx)
    with pytest.raises(pkg.MGVIB200Error, match="NewtonCG"):
        pkg.mgvi_step(m, cfg["data"], 2, cfg["center"], pkg.MGVIConfig(optimizer="LBFGS"), emu_ctx)
    # metric_apply before linearisation is an error, with a message
    V = np.zeros((16, 1), order="F")
    rc = emu_ctx.lib.metric_apply(m.handle, pkg.capi.ptr(V), 1, pkg.capi.ptr(V))
    assert rc != 0 and b"linearize" in emu_ctx.lib.last_error(emu_ctx.handle)
    # bad descriptors
    with pytest.raises((pkg.MGVIB200Error, ValueError)):
        pkg.CorrelatedFieldModel((16,), np.ones(16), 1.0, np.ones(8), emu_ctx)
    with pytest.raises(pkg.MGVIB200Error):
        pkg.CorrelatedFieldModel((16, 2, 2, 2), np.ones(128), 1.0, np.ones(128), emu_ctx)
    with pytest.raises(pkg.MGVIB200Error):
        pkg.CorrelatedFieldModel((16,), np.ones(16), 1.0, np.ones(16), emu_ctx, likelihood=syn.LIKE_NORMAL, sigma=0.0)
    m.close()


def test_config_defaults_match_reference(pkg):
    cfg = pkg.MGVIConfig()
    assert isinstance(cfg.linsolver, pkg.B200CG) and isinstance(cfg.optimizer, pkg.NewtonCG)
    assert cfg.linsolver.abstol == cfg.linsolver.reltol == float(np.sqrt(np.finfo(float).eps))
    o = cfg.optimizer                                   # src/newtoncg.jl:14-31
    assert (o.alpha, o.steps, o.i0, o.i1) == (0.1, 4, 5, 50)
    assert pkg.mgvi_kl_optimize_step is pkg.mgvi_step


def _balanced(src, start):
    """text of the parenthesised group opening at src[start] == '(' (without the outer parens)"""
    depth = 0
    for i in range(start, len(src)):
        if src[i] == "(":
            depth += 1
        elif src[i] == ")":
            depth -= 1
            if depth == 0:
                return src[start + 1:i]
    raise ValueError("unbalanced")


def _top_level_count(text):
    text = text.strip()
    if not text or text == "void":
        return 0
    depth, n = 0, 1
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        elif ch == "," and depth == 0:
            n += 1
    return n - (1 if text.endswith(",") else 0)


def test_julia_shim_matches_the_header():
    """julia/MGVIB200.jl cannot be executed here (no Julia): at least every `ccall` in it must
    name a symbol the header declares, with as many argument types as the C prototype has
    parameters."""
    hdr = open(os.path.join(ROOT, "include", "mgvi_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", "", hdr)
    proto = {}
    for m in re.finditer(r"\b(mgvib200_[a-z0-9_]+)\s*\(", hdr):
        proto[m.group(1)] = _top_level_count(_balanced(hdr, m.end() - 1))
    jl = open(os.path.join(ROOT, "julia", "MGVIB200.jl")).read()
    calls = list(re.finditer(r"ccall\(\(:(mgvib200_[a-z0-9_]+),\s*libmgvib200\),\s*[A-Za-z{}.]+,\s*\(", jl))
    assert len(calls) >= 12
    for m in calls:
        name = m.group(1)
        assert name in proto, "shim calls undeclared symbol %s" % name
        ntypes = _top_level_count(_balanced(jl, m.end() - 1))
        assert ntypes == proto[name], (name, ntypes, proto[name])


def test_model_descriptor_layouts_agree():
    """The model descriptor is declared three times (C header, ctypes mirror, Julia shim): same fields, same
    order -- a field appended in one place only would shift every later member."""
    import __graft_entry__ as g
    pkg = g.load_package()
    py = [f[0] for f in pkg.capi.ModelDesc._fields_]
    hdr = open(os.path.join(ROOT, "include", "mgvi_b200.h")).read()
    body = hdr[hdr.index("typedef struct {\n    int32_t kind;"):hdr.index("} mgvib200_model_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt or stmt.startswith("typedef"):
            stmt = stmt.replace("typedef struct {", "").strip()
            if not stmt:
                continue
        decls = stmt.split(",")
        c.append(re.search(r"([a-z_A-Z0-9]+)(?:\[\d+\])?\s*$", decls[0].strip()).group(1))
        c += [re.search(r"([a-z_A-Z0-9]+)(?:\[\d+\])?\s*$", d.strip()).group(1) for d in decls[1:]]
    assert c == py, (c, py)
    jl = open(os.path.join(ROOT, "julia", "MGVIB200.jl")).read()
    jb = jl[jl.index("struct ModelDesc\n"):]
    jb = jb[:jb.index("\nend")]
    j = [m.group(1) for m in re.finditer(r"^\s+([a-z_A-Z0-9]+)::", jb, flags=re.M)]
    assert j == py, (j, py)
