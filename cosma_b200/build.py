"""Builds libcosma_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the library is a
plain C-ABI shared object, loaded with ctypes / linked by C++ hosts)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libcosma_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nccl_include():
    """nccl.h: prefer the header of the NCCL that torch bundles (the one loaded at run time), else the system one."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            p = os.path.join(list(spec.submodule_search_locations)[0], "include")
            if os.path.exists(os.path.join(p, "nccl.h")):
                return p
    except Exception:
        pass
    return "/usr/include"


NCCL_INC = _nccl_include()


def sources():
    out = []
    for root, _, files in os.walk(CSRC):
        for f in sorted(files):
            if f.endswith((".cu", ".cpp")):
                out.append(os.path.join(root, f))
    return sorted(out)


def headers():
    out = []
    for base in (CSRC, os.path.join(HERE, "..", "include")):
        for root, _, files in os.walk(base):
            out += [os.path.join(root, f) for f in files if f.endswith((".h", ".hpp", ".cuh"))]
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in sources() + headers() + [__file__])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I", os.path.join(HERE, "..", "include"),
              "-I", CSRC, "-I", NCCL_INC, "-DCOSMA_B200_BUILD"]
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.relpath(src, CSRC).replace(os.sep, "_") + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(
                [os.path.getmtime(src)] + [os.path.getmtime(h) for h in headers()] + [os.path.getmtime(__file__)]):
            continue
        cmd = [NVCC] + ARCH + common + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    link = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl", "-lpthread"]
    subprocess.check_call(link)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
