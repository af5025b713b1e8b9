"""Build the CUDA extension in-tree: openvo_b200/lib/libopenvo_b200.so (nvcc, sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libopenvo_b200.so")
CU = ["api.cu", "sgbm.cu", "orb.cu", "match.cu", "filters.cu", "pnp.cu"]
CPP = ["host_select.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--fmad=false",
              "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-Xptxas", "-v"]


def _stale(out, deps):
    return not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps)


def build_variant(name, defines):
    """Experiment helper: a separately named library with extra -D flags (openvo_b200/lib/variants/<name>.so)."""
    vdir = os.path.join(LIBDIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    for f in CU + CPP:
        obj = os.path.join(vdir, name + "_" + f + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-x", "cu"] if f.endswith(".cpp") else []) + ["-c", os.path.join(CSRC, f), "-o", obj]
        subprocess.run(cmd, capture_output=True, text=True, check=True)
        objs.append(obj)
    out = os.path.join(vdir, name + ".so")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + ["-lcudart"])
    for o in objs:
        os.remove(o)
    return out


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "openvo_b200.h"))
    objs = []
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    for f in CU + CPP:
        src = os.path.join(CSRC, f)
        obj = os.path.join(LIBDIR, f + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + (["-x", "cu"] if f.endswith(".cpp") else []) + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            with open(obj + ".log", "w") as lf:
                lf.write(r.stdout + r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed on " + f)
            if verbose:
                print(r.stderr)
    if force or _stale(LIB, objs):
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
