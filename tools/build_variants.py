"""Tuning helper: build libmgvib200 variants that differ in the resident-blocks-per-SM the pass
kernels are compiled for (-DMGVI_MINB=k).  Output: build/variants/libmgvib200_mb<k>.so"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main(minbs):
    g.build_cuda()
    out = os.path.join(g.BUILD, "variants")
    os.makedirs(out, exist_ok=True)

    def one(mb):
        obj = os.path.join(out, "field_engine_mb%d.o" % mb)
        subprocess.check_call(["nvcc"] + g.NVCC_FLAGS + ["-DMGVI_MINB=%d" % mb, "-I", g.CSRC, "-c",
                               os.path.join(g.CSRC, "field_engine.cu"), "-o", obj],
                              stdout=open(obj + ".log", "w"), stderr=subprocess.STDOUT)
        objs = [os.path.join(g.BUILD, s + ".o") for s in g.CU_SOURCES if s != "field_engine"] + [obj]
        lib = os.path.join(out, "libmgvib200_mb%d.so" % mb)
        subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs +
                              ["-lcudart", "-ldl"])
        return lib

    with ThreadPoolExecutor(max_workers=len(minbs)) as ex:
        for lib in ex.map(one, minbs):
            print(lib)


if __name__ == "__main__":
    main([int(a) for a in sys.argv[1:]] or [2, 3, 4])
