"""Tuning helper: build libmgvib200 variants that differ in compile-time switches of the pass
kernels (field_engine.cu is the only translation unit that sees them).

    python tools/build_variants.py 2 3 4                      # -DMGVI_MINB=k  -> build/variants/libmgvib200_mb<k>.so
    python tools/build_variants.py --define MGVI_TW_SQUARE --tag twsq
    python tools/build_variants.py --define MGVI_TW_SQUARE --tag twsq --emu   # emulator build of the same variant

Use a variant with MGVIB200_LIB=<path> (bench.py, tests)."""
import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def build_cuda_variant(defines, tag):
    g.build_cuda()
    out = os.path.join(g.BUILD, "variants")
    os.makedirs(out, exist_ok=True)
    obj = os.path.join(out, "field_engine_%s.o" % tag)
    subprocess.check_call(["nvcc"] + g.NVCC_FLAGS + ["-D" + d for d in defines] +
                          ["-I", g.CSRC, "-c", os.path.join(g.CSRC, "field_engine.cu"), "-o", obj],
                          stdout=open(obj + ".log", "w"), stderr=subprocess.STDOUT)
    objs = [os.path.join(g.BUILD, s + ".o") for s in g.CU_SOURCES if s != "field_engine"] + [obj]
    lib = os.path.join(out, "libmgvib200_%s.so" % tag)
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs +
                          ["-lcudart", "-ldl"])
    return lib


def build_emu_variant(defines, tag):
    """TEST INFRASTRUCTURE: the emulator build (tests/simt_emu) of the same variant."""
    g.build_emu()
    out = os.path.join(g.BUILD, "variants")
    os.makedirs(out, exist_ok=True)
    obj = os.path.join(out, "emu_field_engine_%s.o" % tag)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-DMGVI_SIMT_EMU"] + ["-D" + d for d in defines] +
                          ["-x", "c++", "-I", g.EMU_DIR, "-I", g.CSRC, "-c", os.path.join(g.CSRC, "field_engine.cu"),
                           "-o", obj])
    objs = [os.path.join(g.BUILD, "emu_" + s + ".o") for s in g.CU_SOURCES if s != "field_engine"] + [obj]
    lib = os.path.join(out, "libmgvib200_emu_%s.so" % tag)
    subprocess.check_call(["g++", "-shared", "-o", lib] + objs)
    return lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("minb", nargs="*", type=int, help="resident blocks per SM variants (-DMGVI_MINB=k)")
    ap.add_argument("--define", action="append", default=[], help="extra -D define (repeatable)")
    ap.add_argument("--tag", default=None)
    ap.add_argument("--emu", action="store_true", help="build the emulator variant instead of the CUDA one")
    a = ap.parse_args()
    if a.define:
        tag = a.tag or "_".join(d.split("=")[0].lower() for d in a.define)
        print((build_emu_variant if a.emu else build_cuda_variant)(a.define, tag))
        return
    minbs = a.minb or [2, 3, 4]
    with ThreadPoolExecutor(max_workers=len(minbs)) as ex:
        for lib in ex.map(lambda mb: build_cuda_variant(["MGVI_MINB=%d" % mb], "mb%d" % mb), minbs):
            print(lib)


if __name__ == "__main__":
    main()
