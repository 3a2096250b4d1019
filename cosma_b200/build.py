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


API = os.path.join(CSRC, "api")
LIBDIR = os.path.join(HERE, "lib")
CXX = os.environ.get("CXX", "g++")


def sources():
    """Sources of libcosma_b200.so (the CUDA library behind the C ABI): everything under csrc/ except api/, which is the
    C++ host layer built by g++ into libcosma.so and friends (build_host)."""
    out = []
    for root, _, files in os.walk(CSRC):
        if os.path.abspath(root).startswith(os.path.abspath(API)):
            continue
        for f in sorted(files):
            if f.endswith((".cu", ".cpp")):
                out.append(os.path.join(root, f))
    return sorted(out)


# The C++ host layer (the reference's public API: cosma::multiply, CosmaMatrix, costa::transform, the C interface, p?gemm).
# Plain g++, no CUDA headers: it reaches the GPU only through include/cosma_b200.h. Library split as in the reference
# (src/cosma/CMakeLists.txt:30-110): cosma | cosma_pxgemm_cpp | cosma_pxgemm (ScaLAPACK names) | cosma_prefixed_pxgemm.
HOST_PLANNING = ["strategy.cpp", "mapper.cpp", "interval.cpp", "math_utils.cpp", "environment_variables.cpp", "costa_layout.cpp", "costa_reorder.cpp", "adapt_strategy.cpp", "auto_strategy.cpp", "schedule.cpp"]
HOST_LIBS = [
    ("libcosma.so", [os.path.join("host", f) for f in HOST_PLANNING] +
     [os.path.join("api", f) for f in ("process_group.cpp", "runtime.cpp", "context.cpp", "matrix.cpp", "multiply.cpp", "costa_api.cpp",
                                       "cinterface.cpp")], ["cosma_b200"]),
    ("libcosma_blacs_lite.so", [os.path.join("api", "blacs_lite.cpp")], ["cosma"]),
    ("libcosma_pxgemm_cpp.so", [os.path.join("api", f) for f in ("scalapack.cpp", "cosma_pxgemm.cpp", "costa_pxtransform.cpp")],
     ["cosma", "cosma_b200", "cosma_blacs_lite"]),
    ("libcosta_scalapack.so", [os.path.join("api", "costa_scalapack.cpp")], ["cosma_pxgemm_cpp", "cosma"]),
    ("libcosta_prefixed_scalapack.so", [os.path.join("api", "costa_prefixed_scalapack.cpp")], ["cosma_pxgemm_cpp", "cosma"]),
    ("libcosma_pxgemm.so", [os.path.join("api", "pxgemm.cpp")], ["cosma_pxgemm_cpp", "cosma"]),
    ("libcosma_prefixed_pxgemm.so", [os.path.join("api", "prefixed_pxgemm.cpp")], ["cosma_pxgemm_cpp", "cosma"]),
]


def host_needs_build():
    deps = headers() + [__file__] + [os.path.join(API, f) for f in os.listdir(API)]
    for name, srcs, _ in HOST_LIBS:
        lib = os.path.join(LIBDIR, name)
        if not os.path.exists(lib):
            return True
        t = os.path.getmtime(lib)
        if any(os.path.getmtime(p) > t for p in deps + [os.path.join(CSRC, s) for s in srcs]):
            return True
    return False


def build_host(force=False):
    """g++ build of the C++ host layer. Needs libcosma_b200.so (build()) for linking."""
    if not force and not host_needs_build():
        return [os.path.join(LIBDIR, n) for n, _, _ in HOST_LIBS]
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build", "host")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-Wno-unused-function", "-I", os.path.join(HERE, "..", "include"), "-I", API]
    procs, objs_of = [], {}
    seen = {}
    for name, srcs, _ in HOST_LIBS:
        objs_of[name] = []
        for s in srcs:
            obj = os.path.join(objdir, s.replace(os.sep, "_") + ".o")
            objs_of[name].append(obj)
            if s in seen:
                continue
            seen[s] = obj
            procs.append((s, subprocess.Popen([CXX] + flags + ["-c", os.path.join(CSRC, s), "-o", obj], stdout=subprocess.PIPE,
                                              stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("g++ failed on %s" % s)
        if out.strip():
            sys.stderr.write(out)
    outs = []
    for name, _, libs in HOST_LIBS:
        lib = os.path.join(LIBDIR, name)
        cmd = [CXX, "-shared", "-o", lib] + objs_of[name] + ["-L", LIBDIR] + ["-l" + l for l in libs] + ["-Wl,-rpath,$ORIGIN", "-lpthread", "-ldl"]
        subprocess.check_call(cmd)
        outs.append(lib)
    return outs


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
    for lib in build_host(force="--force" in sys.argv):
        print(lib)
