"""TEST INFRASTRUCTURE ONLY: compile openvo_b200/csrc/*.cu with g++ against tests/emu/cuda_emu.h so the kernels' logic can be
exercised in the CPU-only tier.  The product never loads this library (openvo_b200/_native.py only loads the nvcc build).
Built with -DOVO_BOUNDS: the kernels' own offset checks (OVO_DEVCHECK, csrc/common.cuh) abort the test process when violated."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
CSRC = os.path.join(ROOT, "openvo_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libopenvo_b200_emu.so")
SRCS = ["api.cu", "sgbm.cu", "orb.cu", "match.cu", "filters.cu", "pnp.cu", "host_select.cpp"]


def build(force=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cuda_emu.h"),
                                                                os.path.join(ROOT, "include", "openvo_b200.h")]
    if force or not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        objs = []
        for f in SRCS:
            obj = os.path.join(HERE, "_build", f + ".o")
            src = os.path.join(CSRC, f)
            if force or not os.path.exists(obj) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in deps):
                subprocess.check_call(["g++", "-std=c++17", "-O2", "-g", "-DOVO_EMU", "-I", HERE, "-fPIC", "-mfma",
                                       "-ffp-contract=off", "-Wno-unused-function"] + os.environ.get("OVO_EMU_DEFS", "-DOVO_BOUNDS").split() + [ "-x", "c++", "-c", src, "-o", obj])
            objs.append(obj)
        subprocess.check_call(["g++", "-shared", "-o", OUT] + objs + ["-lpthread"])
    return OUT


if __name__ == "__main__":
    print(build())
