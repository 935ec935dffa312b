"""Build the product shared library in-tree: gpboost_b200/lib_gpboost_b200.so (sm_100a only).

nvcc compiles the CUDA device engine (csrc/dev) and the C++ host layer (csrc/host: REModel, L-BFGS,
GPB_*/LGBM_* C API); g++ is only the host compiler behind nvcc. No torch / pybind: the library is plain C ABI.
"""
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "lib_gpboost_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def lib_path():
    return os.path.join(_HERE, LIB_NAME)


def _sources():
    cu = sorted(glob.glob(os.path.join(_HERE, "csrc", "dev", "*.cu")))
    cpp = sorted(glob.glob(os.path.join(_HERE, "csrc", "host", "*.cpp")))
    hdr = (glob.glob(os.path.join(_HERE, "csrc", "*", "*.cuh")) + glob.glob(os.path.join(_HERE, "csrc", "*", "*.h")) +
           glob.glob(os.path.join(os.path.dirname(_HERE), "include", "*.h")))
    return cu, cpp, hdr


def build(force=False, verbose=False, extra_flags=(), out_name=None):
    """extra_flags / out_name: build an experimental variant next to the product library (tuning runs)."""
    cu, cpp, hdr = _sources()
    out = lib_path() if out_name is None else os.path.join(_HERE, out_name)
    deps = cu + cpp + hdr + [os.path.abspath(__file__)]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in deps):
        return out
    objdir = os.path.join(_HERE, "build" if out_name is None else "build_" + out_name.replace(".", "_"))
    os.makedirs(objdir, exist_ok=True)
    objs = []
    common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fopenmp,-fvisibility=hidden",
              "-I", os.path.join(os.path.dirname(_HERE), "include")] + list(extra_flags)
    procs = []
    for s in cu + cpp:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        if not force and os.path.exists(o) and all(os.path.getmtime(o) >= os.path.getmtime(x) for x in [s] + hdr):
            continue
        cmd = [NVCC] + ARCH + common + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu" if s.endswith(".cu") else "c++", "-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for s, p in procs:
        outp, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (s, outp.decode()))
        if verbose:
            print(outp.decode())
    cmd = [NVCC] + ARCH + ["-shared", "-Xcompiler", "-fPIC,-fopenmp", "-o", out] + objs + ["-lcudart", "-lgomp"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(verbose=False))
