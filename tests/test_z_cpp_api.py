"""The C++ host layer (libcosma.so and friends: cosma::multiply / CosmaMatrix / multiply_using_layout, costa::transform, the C
interface, the ScaLAPACK p?gemm symbols) driven by C++ test programs that read like the reference's own tests
(tests/cpp/*.cpp <-> reference tests/multiply.cpp, scalar_matmul.cpp, multiply_using_layout.cpp, pdgemm.cpp).

CPU part: the libraries build with plain g++, export the reference's symbols, the MPI-name subset works across processes,
the coordinate maps match the reference's goldens, and compute entry points fail loudly without a GPU.
GPU part (-m gpu): the programs run on 1 rank and, when the box has the GPUs, on 2 / 4 / 8 ranks (one per GPU)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CPP = os.path.join(HERE, "cpp")
BIN = os.path.join(CPP, "bin")
LIBDIR = os.path.join(ROOT, "cosma_b200", "lib")
sys.path.insert(0, ROOT)

PROGRAMS = {
    # name: (source, extra libraries, needs the oracle)
    "test_api_cpu": (os.path.join(CPP, "test_api_cpu.cpp"), [], False),
    "test_process_group": (os.path.join(CPP, "test_process_group.cpp"), [], False),
    "test_multiply": (os.path.join(CPP, "test_multiply.cpp"), [], True),
    "test_multiply_using_layout": (os.path.join(CPP, "test_multiply_using_layout.cpp"), [], True),
    "test_costa_examples": (os.path.join(CPP, "test_costa_examples.cpp"), [], False),
    "test_pxtran": (os.path.join(CPP, "test_pxtran.cpp"), ["costa_prefixed_scalapack", "costa_scalapack", "cosma_pxgemm_cpp", "cosma_blacs_lite"], False),
    "test_pxgemm": (os.path.join(CPP, "test_pxgemm.cpp"), ["cosma_prefixed_pxgemm", "cosma_pxgemm", "cosma_pxgemm_cpp", "cosma_blacs_lite"], True),
    "cosma_miniapp": (os.path.join(ROOT, "miniapp", "cosma_miniapp.cpp"), [], False),
    "pxgemm_miniapp": (os.path.join(ROOT, "miniapp", "pxgemm_miniapp.cpp"), ["cosma_pxgemm_cpp", "cosma_blacs_lite"], False),
    "pxgemr2d_miniapp": (os.path.join(ROOT, "miniapp", "pxgemr2d_miniapp.cpp"), ["cosma_pxgemm_cpp", "cosma_blacs_lite"], False),
    "pxtran_miniapp": (os.path.join(ROOT, "miniapp", "pxtran_miniapp.cpp"), ["cosma_pxgemm_cpp", "cosma_blacs_lite"], False),
}


@pytest.fixture(scope="session")
def host_libs(lib):
    from cosma_b200 import build
    return build.build_host()


def program(name, oracle=None):
    src, libs, needs_oracle = PROGRAMS[name]
    out = os.path.join(BIN, name)
    os.makedirs(BIN, exist_ok=True)
    deps = [src] + [os.path.join(d, f) for d in (CPP, os.path.join(ROOT, "miniapp")) for f in os.listdir(d) if f.endswith(".hpp")]
    deps.append(os.path.join(LIBDIR, "libcosma.so"))
    if os.path.exists(out) and all(os.path.getmtime(out) > os.path.getmtime(d) for d in deps):
        return out
    cmd = ["g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", out, "-L", LIBDIR]
    cmd += ["-l" + l for l in libs] + ["-lcosma", "-lcosma_b200", "-Wl,-rpath," + LIBDIR]
    if needs_oracle:
        odir = os.path.join(ROOT, "oracle")
        cmd += ["-L", odir, "-loracle", "-Wl,-rpath," + odir]
    subprocess.check_call(cmd)
    return out


def run_ranks(np_, argv, timeout=180):
    from cosma_b200.launch import launch
    code, outs = launch(np_, argv, timeout=timeout, capture=True)
    text = "\n".join("--- rank %d ---\n%s" % (r, o) for r, o in enumerate(outs))
    assert code == 0, text[-6000:]
    return outs[0]


# ---- CPU --------------------------------------------------------------------------------------------------------------

def test_host_libraries_export_the_reference_symbols(host_libs):
    def symbols(name):
        out = subprocess.check_output(["nm", "-D", "--defined-only", "-C", os.path.join(LIBDIR, name)], text=True)
        while "> >" in out:
            out = out.replace("> >", ">>")
        return out
    core = symbols("libcosma.so")
    for t in ("float", "double", "std::complex<float>", "std::complex<double>"):
        assert "void cosma::multiply<%s>(cosma::CosmaMatrix<%s>&" % (t, t) in core, t
        assert "void cosma::multiply_using_layout<%s>(costa::grid_layout<%s>&" % (t, t) in core, t
        assert "cosma::CosmaMatrix<%s>::matrix_size() const" % t in core, t
        assert "cosma::get_context_instance<%s>()" % t in core, t
        assert "void costa::transform<%s>(costa::grid_layout<%s>&, costa::grid_layout<%s>&, char, %s, %s" % (t, t, t, t, t) in core, t
    for f in ("smultiply_using_layout", "dmultiply_using_layout", "cmultiply_using_layout", "zmultiply_using_layout"):
        assert " T %s\n" % f in core, f
    assert "cosma::Strategy::Strategy(int, int, int, unsigned long, long long, bool, bool, bool)" in core
    px = symbols("libcosma_pxgemm.so")
    pre = symbols("libcosma_prefixed_pxgemm.so")
    for t in "sdcz":
        for name in ("p%sgemm" % t, "p%sgemm_" % t, ("p%sgemm" % t).upper(), ("p%sgemm_" % t).upper()):
            assert " T %s\n" % name in px, name
        for name in ("cosma_p%sgemm" % t, "cosma_p%sgemm_" % t, "COSMA_P%sGEMM" % t.upper(), "COSMA_P%sGEMM_" % t.upper()):
            assert " T %s\n" % name in pre, name
    sc, psc = symbols("libcosta_scalapack.so"), symbols("libcosta_prefixed_scalapack.so")
    for low in ["p%sgemr2d" % t for t in "sdcz"] + ["pstran", "pdtran", "pctranu", "pztranu", "pctranc", "pztranc"]:
        for name in (low, low + "_", low.upper(), low.upper() + "_"):
            assert " T %s\n" % name in sc, name
        for name in ("costa_" + low, "costa_" + low + "_", "COSTA_" + low.upper(), "COSTA_" + low.upper() + "_"):
            assert " T %s\n" % name in psc, name
    bl = symbols("libcosma_blacs_lite.so")
    for name in ("Cblacs_gridinit", "Cblacs_gridinfo", "Cblacs2sys_handle", "Cblacs_pcoord", "blacs_gridinit_", "blacs_gridinfo_", "blacs_pinfo_",
                 "blacs_get_", "blacs_pnum_", "blacs_pcoord_", "blacs_barrier_", "blacs_gridexit_", "blacs_exit_", "descinit_", "numroc_"):
        assert " W %s\n" % name in bl, name   # weak: a real BLACS takes precedence
    cpp = symbols("libcosma_pxgemm_cpp.so")
    assert "void costa::pxgemr2d<double>(int, int, double const*" in cpp and "void costa::pxtran_op<std::complex<float>>(" in cpp
    assert "void cosma::pxgemm<double>(char, char, int, int, int, double, double const*" in cpp
    assert "void cosma::pxgemm<std::complex<double>>(" in cpp


@pytest.mark.parametrize("header", ["cosma/multiply.hpp", "cosma/matrix.hpp", "cosma/context.hpp", "cosma/strategy.hpp", "cosma/mapper.hpp",
                                    "cosma/cinterface.hpp", "cosma/cosma_pxgemm.hpp", "cosma/pxgemm.h", "cosma/prefixed_pxgemm.h",
                                    "costa/layout.hpp", "costa/grid2grid/transform.hpp", "costa/grid2grid/transformer.hpp",
                                    "costa/grid2grid/scalapack_layout.hpp", "costa/grid2grid/comm_volume.hpp", "costa/grid2grid/ranks_reordering.hpp",
                                    "costa/pxgemr2d/costa_pxgemr2d.hpp", "costa/pxtran_op/costa_pxtran_op.hpp", "costa/pxgemr2d/pxgemr2d.h",
                                    "costa/pxtran/pxtran.h", "costa/pxtranu/prefixed_pxtranu.h", "costa/pxtranc/pxtranc.h", "cosma/blacs.hpp",
                                    "cosma/scalapack.hpp", "cosma/process_group.hpp", "cosma_b200.h"])
def test_public_headers_are_self_contained(header, tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text("#include <%s>\nint main() { return 0; }\n" % header)
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), str(src)])


@pytest.mark.parametrize("source", ["runtime", "context", "matrix", "multiply", "costa_api", "cinterface", "scalapack", "cosma_pxgemm", "costa_pxtransform",
                                    "pxgemm", "prefixed_pxgemm", "costa_scalapack", "costa_prefixed_scalapack"])
def test_host_layer_compiles_against_a_real_mpi_header(source):
    """-DCOSMA_B200_WITH_MPI: MPI_Comm and the MPI calls come from <mpi.h> instead of mpi_compat's process group. No MPI is installed
    here; the declaration-only mpi.h of the reference checker (oracle/stubs/, MPI_Comm = int) stands in for the header, which is enough
    to prove that the host layer only uses MPI names a real MPI declares (MPI_Comm_c2f is supplied on the command line)."""
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-DCOSMA_B200_WITH_MPI", "-DMPI_Comm_c2f(c)=(c)", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "cosma_b200", "csrc", "api"), "-I", os.path.join(ROOT, "oracle", "stubs"),
                           os.path.join(ROOT, "cosma_b200", "csrc", "api", source + ".cpp")])


@pytest.mark.parametrize("source", ["tests/cpp/test_multiply.cpp", "tests/cpp/test_multiply_using_layout.cpp", "tests/cpp/test_pxgemm.cpp",
                                    "tests/cpp/test_pxtran.cpp", "tests/cpp/test_costa_examples.cpp", "miniapp/cosma_miniapp.cpp", "miniapp/pxgemm_miniapp.cpp",
                                    "miniapp/pxgemr2d_miniapp.cpp", "miniapp/pxtran_miniapp.cpp"])
def test_programs_compile_against_a_real_mpi_header(source):
    """The test programs and miniapps only use MPI calls that exist in MPI: they compile unchanged with -DCOSMA_B200_WITH_MPI."""
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-DCOSMA_B200_WITH_MPI", "-DMPI_Comm_c2f(c)=(c)", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, source)])


def test_host_layer_runs_over_an_mpi_implementation(host_libs, oracle):
    """The -DCOSMA_B200_WITH_MPI build AT RUN TIME: api/ + host/ + tests/cpp/test_multiply.cpp compiled against the MPI header and linked
    with the multi-process MPI stand-in that runs the unmodified reference (oracle/stubs/minimpi.cpp, started by oracle/minirun.py),
    the C ABI underneath replaced by the CPU stand-in. 4 ranks: MPI_Comm_create_group of the active ranks, the ncclUniqueId-style
    broadcast, idle ranks, all multiply cases that fit 4 ranks."""
    import socket
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from minirun import launch as mpi_launch
    api, host = os.path.join(ROOT, "cosma_b200", "csrc", "api"), os.path.join(ROOT, "cosma_b200", "csrc", "host")
    stubs = os.path.join(ROOT, "oracle", "stubs")
    exe = os.path.join(BIN, "mpi_test_multiply")
    srcs = [os.path.join(api, f) for f in ("process_group.cpp", "runtime.cpp", "context.cpp", "matrix.cpp", "multiply.cpp", "costa_api.cpp", "cinterface.cpp")]
    srcs += [os.path.join(host, f) for f in ("strategy.cpp", "mapper.cpp", "interval.cpp", "math_utils.cpp", "environment_variables.cpp", "costa_layout.cpp",
                                             "costa_reorder.cpp")]
    os.makedirs(BIN, exist_ok=True)
    mock_o = os.path.join(BIN, "mock_b200_nompi.o")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-c", os.path.join(CPP, "mock_b200.cpp"), "-I", os.path.join(ROOT, "include"), "-o", mock_o])
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DCOSMA_B200_WITH_MPI", "-DMPI_Comm_c2f(c)=(c)", "-I", os.path.join(ROOT, "include"), "-I", api, "-I", stubs,
                           os.path.join(CPP, "test_multiply.cpp")] + srcs + [os.path.join(stubs, "minimpi.cpp"), mock_o, "-o", exe,
                           "-L", os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, COSMA_B200_PG_PORT=str(port), MASTER_ADDR="127.0.0.1")
    code, outs = mpi_launch(4, ["bash", "-c", "RANK=$MINIMPI_RANK WORLD_SIZE=$MINIMPI_SIZE exec " + exe], env=env, stdout=subprocess.PIPE, timeout=300)
    text = outs[0].decode()
    assert code == 0 and "failed = 0" in text and "idle ranks" in text, text[-3000:]


def test_c_abi_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text("#include <cosma_b200.h>\nint main(void) { return cosma_b200_version() == 0; }\n")
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), str(src)])


def test_api_on_cpu(host_libs):
    out = subprocess.run([program("test_api_cpu")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "failed = 0" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("np_", [1, 2, 5])
def test_process_group(host_libs, np_):
    out = run_ranks(np_, [program("test_process_group")], timeout=120)
    assert "failed = 0" in out, out


def test_compute_fails_loudly_without_gpu(host_libs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    exe = program("cosma_miniapp")
    out = subprocess.run([exe, "-m", "64", "-n", "64", "-k", "64"], capture_output=True, text=True, timeout=120)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stdout + out.stderr) or "cuda" in (out.stdout + out.stderr).lower()


def _mock():
    """tests/cpp/mock_b200.cpp -> libmock_b200.so: CPU stand-in for the C ABI entry points the host layer calls (test infrastructure)."""
    src, out = os.path.join(CPP, "mock_b200.cpp"), os.path.join(BIN, "libmock_b200.so")
    os.makedirs(BIN, exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(LIBDIR, "libcosma.so"))):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), src, "-o", out,
                               "-L", LIBDIR, "-lcosma", "-Wl,-rpath," + LIBDIR])
    return out


def _run_on_mock(np_, argv, timeout=300):
    from cosma_b200.launch import launch
    code, outs = launch(np_, argv, timeout=timeout, capture=True, env_extra={"LD_PRELOAD": _mock()})
    text = "\n".join("--- rank %d ---\n%s" % (r, o) for r, o in enumerate(outs))
    assert code == 0, text[-6000:]
    return outs[0]


@pytest.mark.parametrize("name,np_", [("test_multiply", 2), ("test_multiply", 4), ("test_multiply", 7), ("test_multiply", 16),
                                      ("test_multiply_using_layout", 2), ("test_multiply_using_layout", 6), ("test_pxgemm", 2), ("test_pxgemm", 6),
                                      ("test_costa_examples", 4), ("test_pxtran", 1), ("test_pxtran", 4), ("test_pxtran", 6)])
def test_cpp_programs_multirank_on_cpu(host_libs, oracle, name, np_):
    """The C++ test programs on 2..16 RANKS without a GPU: the whole host layer is real (communicators, idle ranks, strategies,
    coordinate maps, layout conversion, BLACS-lite, the MPI-name subset, the programs' own message protocols); only the C ABI entry
    points underneath are replaced by a gather -> naive GEMM / dense relayout -> scatter stand-in (tests/cpp/mock_b200.cpp). At 16
    ranks all 40 cases of the reference's tests/multiply.cpp run. Guards the GPU budget against protocol deadlocks."""
    out = _run_on_mock(np_, [program(name)])
    assert "failed = 0" in out, out[-4000:]
    assert "checks passed (all ranks) = 0," not in out


def test_dim_threshold_hands_small_problems_to_the_next_scalapack(host_libs):
    """COSMA_DIM_THRESHOLD: libcosma_pxgemm.so in front of a (fake) ScaLAPACK; small problems reach the fake, large ones are served."""
    fake = os.path.join(BIN, "libfake_scalapack.so")
    os.makedirs(BIN, exist_ok=True)
    subprocess.check_call(["gcc", "-O1", "-fPIC", "-shared", os.path.join(CPP, "fake_scalapack.c"), "-o", fake])
    exe = os.path.join(BIN, "test_interpose")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(CPP, "test_interpose.cpp"), "-o", exe,
                           "-L", LIBDIR, "-L", BIN, "-lcosma_pxgemm", "-Wl,--no-as-needed", "-lfake_scalapack", "-Wl,--as-needed", "-lcosma_pxgemm_cpp", "-lcosma_blacs_lite", "-lcosma",
                           "-lcosma_b200", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + BIN])
    from cosma_b200.launch import launch
    code, outs = launch(1, [exe], timeout=120, capture=True, env_extra={"LD_PRELOAD": _mock(), "COSMA_DIM_THRESHOLD": "64"})
    assert code == 0 and "failed = 0" in outs[0], outs[0]


@pytest.mark.parametrize("np_,args", [(4, ["-m", "600", "-n", "500", "-k", "700", "-r", "2"]),
                                      (5, ["-m", "300", "-n", "300", "-k", "300", "-r", "1", "-t", "zdouble"]),   # the strategy idles ranks
                                      (6, ["-m", "640", "-n", "640", "-k", "640", "-s", "pm2,pk3", "-r", "1", "-t", "float"])])
def test_cosma_miniapp_multirank_on_cpu(host_libs, np_, args):
    out = _run_on_mock(np_, [program("cosma_miniapp")] + args)
    assert "COSMA TIMES [ms] =" in out and "Strategy" in out, out


def test_numa_binding_switch_is_best_effort(host_libs):
    """COSMA_B200_BIND_NUMA=ON: the runtime reads the local_cpulist of the device's PCI function and restricts the process to it when
    that is a proper subset of its CPUs; here (one NUMA node, or no such sysfs entry) nothing changes and the run is unaffected."""
    from cosma_b200.launch import launch
    devs = sorted(os.listdir("/sys/bus/pci/devices")) if os.path.isdir("/sys/bus/pci/devices") else []
    for bdf in ([devs[0]] if devs else []) + ["ffff:ff:1f.7"]:
        code, outs = launch(2, [program("cosma_miniapp"), "-m", "96", "-n", "80", "-k", "64", "-r", "1"], timeout=120, capture=True,
                            env_extra={"LD_PRELOAD": _mock(), "COSMA_B200_BIND_NUMA": "ON", "MOCK_PCI_BDF": bdf, "COSMA_B200_TRACE": "ON"})
        assert code == 0 and "COSMA TIMES [ms] =" in outs[0], outs[0][-2000:]


def test_pxgemm_miniapp_multirank_on_cpu(host_libs):
    out = _run_on_mock(6, [program("pxgemm_miniapp"), "-m", "200", "-n", "150", "-k", "100", "--block_a", "32,16", "--block_b", "8,8", "--block_c", "16,32",
                           "--transpose", "TN", "-p", "2,3", "-r", "2", "--type", "zdouble"])
    assert "COSMA TIMES [ms] =" in out and "grid 2 x 3" in out, out


@pytest.mark.parametrize("np_,args", [
    (1, ["-m", "100", "-n", "77", "--block_a", "16,8", "--block_c", "5,32"]),
    (6, ["-m", "301", "-n", "203", "--block_a", "32,16", "--block_c", "7,50", "-p", "2,3", "-q", "3,2", "-t", "zdouble"]),
    (6, ["-m", "128", "-n", "256", "--block_a", "32,32", "--block_c", "32,32", "-p", "2,2", "-q", "1,6", "-t", "float"]),   # grid A leaves ranks out
    (4, ["-m", "90", "-n", "90", "--block_a", "9,9", "--block_c", "10,10", "-p", "3,3", "-t", "zfloat"]),                 # wrong grid -> 1 x P
])
def test_pxgemr2d_miniapp_multirank_on_cpu(host_libs, np_, args):
    """libs/COSTA/miniapps/pxgemr2d_miniapp.cpp's counterpart with --test: every rank checks its part of C against the definition."""
    out = _run_on_mock(np_, [program("pxgemr2d_miniapp")] + args + ["--test"])
    assert "COSTA TIMES [ms] =" in out and "Result is CORRECT!" in out, out


@pytest.mark.parametrize("np_,args", [
    (1, ["-m", "60", "-n", "45", "--block_a", "8,8", "--block_c", "16,4"]),
    (6, ["-m", "301", "-n", "203", "--block_a", "32,16", "--block_c", "7,50", "-p", "2,3", "-t", "zdouble", "--op", "C", "--alpha", "2", "--beta", "-1"]),
    (4, ["-m", "128", "-n", "64", "--block_a", "32,32", "--block_c", "32,32", "-t", "float", "--alpha", "3"]),
    (4, ["-m", "50", "-n", "70", "--block_a", "3,5", "--block_c", "4,6", "-t", "zfloat", "--beta", "1"]),
])
def test_pxtran_miniapp_multirank_on_cpu(host_libs, np_, args):
    """libs/COSTA/miniapps/pxtran_miniapp.cpp's counterpart with --test (transpose / conjugate transpose, integer alpha and beta)."""
    out = _run_on_mock(np_, [program("pxtran_miniapp")] + args + ["--test"])
    assert "COSTA TIMES [ms] =" in out and "Result is CORRECT!" in out, out


# ---- GPU --------------------------------------------------------------------------------------------------------------

def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _skip_unless_ranks(np_):
    """Multi-rank runs need as many GPUs (one rank per GPU). Seen green on hardware: 2 ranks (profiles/r2b_pytest_gpu_n2.txt)."""
    if np_ > 1 and np_ > _gpus():
        pytest.skip("needs %d GPUs" % np_)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["test_multiply", "test_multiply_using_layout", "test_pxgemm", "test_costa_examples", "test_pxtran"])
@pytest.mark.parametrize("np_", [1, 2, 4, 8])
def test_cpp_program(host_libs, oracle, name, np_):
    _skip_unless_ranks(np_)
    if name == "test_costa_examples" and np_ != 4:
        pytest.skip("the COSTA examples are written for 4 ranks")
    out = run_ranks(np_, [program(name)], timeout=240)
    assert "failed = 0" in out, out[-4000:]
    assert "checks passed (all ranks) = 0," not in out, out[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("np_", [1, 2, 4, 8])
@pytest.mark.parametrize("dtype", ["double", "zdouble", "float", "zfloat"])
def test_cosma_miniapp(host_libs, dtype, np_):
    _skip_unless_ranks(np_)
    out = run_ranks(np_, [program("cosma_miniapp"), "-m", "1024", "-n", "768", "-k", "1280", "-r", "2", "-t", dtype], timeout=120)
    assert "COSMA TIMES [ms] =" in out, out


@pytest.mark.gpu
@pytest.mark.parametrize("np_", [1, 2, 4, 8])
def test_pxgemm_miniapp(host_libs, np_):
    _skip_unless_ranks(np_)
    out = run_ranks(np_, [program("pxgemm_miniapp"), "-m", "1024", "-n", "768", "-k", "512", "--block_a", "128,128", "--block_b", "64,64",
                          "--block_c", "128,32", "--trans_a", "T", "-r", "2", "--type", "zdouble"], timeout=120)
    assert "COSMA TIMES [ms] =" in out, out


@pytest.mark.gpu
@pytest.mark.parametrize("np_", [1, 2, 4, 8])
@pytest.mark.parametrize("dtype", ["double", "zfloat"])
def test_costa_miniapps(host_libs, dtype, np_):
    """p?gemr2d and p?tran(c) miniapps with --test: host-resident block-cyclic arrays, exact check on every rank."""
    _skip_unless_ranks(np_)
    out = run_ranks(np_, [program("pxgemr2d_miniapp"), "-m", "1500", "-n", "1100", "--block_a", "128,64", "--block_c", "50,200", "-t", dtype, "--test"], timeout=120)
    assert "Result is CORRECT!" in out, out
    out = run_ranks(np_, [program("pxtran_miniapp"), "-m", "1500", "-n", "1100", "--block_a", "128,64", "--block_c", "50,200", "-t", dtype, "--op", "C",
                          "--alpha", "2", "--beta", "-1", "--test"], timeout=120)
    assert "Result is CORRECT!" in out, out
