"""GPU tests of switches that are still opt-in because they were written after the round's GPU budget was spent. They
run LAST in the suite so that a surprise here cannot hide the established kernel suites under `pytest -x`.

COSMA_B200_REPACK_UNALIGNED=ON: operands the TMA path cannot address (odd leading dimension, 8-byte-aligned base) are repacked once
into a stream-ordered scratch and the tensor-pipe kernel runs on the copies (csrc/repack.h). Must give the same numbers as the
oracle and report the tensor-pipe path (last_gemm_path == 1) where the default build reports the generic kernel (2).

COSMA_B200_CACHED_PROBLEMS=2: the per-communicator cache of p?gemm / multiply_using_layout problems (plan + three device arenas each)
drops the least recently used problem; walking twice through six shapes must keep giving the dense result and must not accumulate
device memory."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
import numpy as np
import torch
sys.path.insert(0, %(root)r)
from cosma_b200 import gemm, _lib
from oracle import oracle as orc
lib = _lib.load()
rng = np.random.default_rng(3)
ok = True
for dt, ta, tb, m, n, k, lda_pad, off in (("d", "N", "N", 301, 257, 199, 0, 0), ("d", "T", "N", 300, 260, 201, 1, 1), ("d", "N", "T", 257, 255, 128, 2, 3),
                                           ("s", "N", "N", 301, 257, 199, 0, 0), ("s", "T", "T", 258, 262, 130, 1, 1)):
    npdt = {"d": np.float64, "s": np.float32}[dt]
    ar, ac = (m, k) if ta == "N" else (k, m)
    br, bc = (k, n) if tb == "N" else (n, k)
    lda, ldb, ldc = ar + lda_pad, br + lda_pad, m + 1
    A = rng.integers(-3, 4, size=lda * ac + off).astype(npdt)
    B = rng.integers(-3, 4, size=ldb * bc + off).astype(npdt)
    C = rng.integers(-3, 4, size=ldc * n).astype(npdt)
    want = orc.gemm(ta, tb, m, n, k, 2.0, A[off:], lda, B[off:], ldb, -1.0, C.copy(), ldc)
    dA, dB, dC = (torch.from_numpy(x).cuda() for x in (A, B, C))
    es = A.itemsize
    gemm.gemm_raw(dt, ta, tb, m, n, k, 2.0, dA.data_ptr() + off * es, lda, dB.data_ptr() + off * es, ldb, -1.0, dC.data_ptr(), ldc)
    torch.cuda.synchronize()
    path = lib.cosma_b200_last_gemm_path()
    same = np.array_equal(dC.cpu().numpy(), want)   # small integers: exact in every type
    print(dt, ta, tb, m, n, k, "path", path, "exact", same)
    ok = ok and same and path == %(path)d
print("RESULT", "OK" if ok else "FAILED")
'''


def _run(env_value, expect_path):
    env = dict(os.environ)
    env.pop("COSMA_B200_REPACK_UNALIGNED", None)
    if env_value:
        env["COSMA_B200_REPACK_UNALIGNED"] = env_value
    out = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "path": expect_path}], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    return out.returncode, out.stdout + out.stderr


@pytest.mark.gpu
def test_unaligned_operands_default_generic_kernel(lib, oracle):
    code, text = _run(None, 2)
    assert code == 0 and "RESULT OK" in text, text[-3000:]


@pytest.mark.gpu
def test_unaligned_operands_repacked_onto_the_tensor_pipe(lib, oracle):
    code, text = _run("ON", 1)
    assert code == 0 and "RESULT OK" in text, text[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("dt", ["d", "s", "z"])
def test_large_unaligned_operands_are_repacked_by_default(lib, oracle, dt):
    """COSMA_B200_REPACK_UNALIGNED unset = AUTO: from m n k >= 2^27 on, operands with an odd leading dimension (COSMA's native layout of an
    irregular split) are repacked by the copy kernel of csrc/repack.cu and multiplied on the tensor pipe (path 1); exact on small integers."""
    import numpy as np
    import torch
    from cosma_b200 import gemm, _lib
    if os.environ.get("COSMA_B200_REPACK_UNALIGNED"):
        pytest.skip("explicit repack mode")
    rng = np.random.default_rng(5)
    m, n, k = 1023, 1024, 1023
    npdt = {"d": np.float64, "s": np.float32, "z": np.complex128}[dt]
    lda, ldb, ldc = m, k, m                      # odd pitches: not addressable by TMA
    off = 1 if dt != "z" else 0                  # and a base that is only element-aligned
    def ints(cnt):
        v = rng.integers(-3, 4, size=cnt).astype(np.float64)
        return (v + 1j * rng.integers(-3, 4, size=cnt)).astype(npdt) if dt == "z" else v.astype(npdt)
    A, B, C = ints(lda * k + off), ints(ldb * n + off), ints(ldc * n)
    dA, dB, dC = (torch.from_numpy(x).cuda() for x in (A, B, C))
    es = A.itemsize
    gemm.gemm_raw(dt, "N", "N", m, n, k, 2.0, dA.data_ptr() + off * es, lda, dB.data_ptr() + off * es, ldb, -1.0, dC.data_ptr(), ldc)
    torch.cuda.synchronize()
    assert _lib.load().cosma_b200_last_gemm_path() == 1
    wide = np.complex128 if dt == "z" else np.float64
    want = 2.0 * (A[off:].reshape(k, lda).T.astype(wide) @ B[off:].reshape(n, ldb).T.astype(wide)) - C.reshape(n, ldc).T.astype(wide)
    assert np.array_equal(dC.cpu().numpy().reshape(n, ldc).T, want.astype(npdt))


CACHE_SCRIPT = r'''
import sys
import torch
sys.path.insert(0, %(root)r)
from cosma_b200 import costa
from cosma_b200.distributed import init_comm
comm = init_comm()
grid = costa.Grid(comm, "R", 1, 1)
gen = torch.Generator(device="cuda").manual_seed(11)


def one(m, n, k):
    # column-major m x k etc. held as the transposed row-major tensors
    At = torch.randint(-4, 5, (k, m), generator=gen, device="cuda").double()
    Bt = torch.randint(-4, 5, (n, k), generator=gen, device="cuda").double()
    Ct = torch.full((n, m), float("nan"), dtype=torch.float64, device="cuda")
    desc = lambda r, c: [1, 0, r, c, 64, 32, 0, 0, r]
    costa.pxgemm(grid, "d", "N", "N", m, n, k, 1.0, At.data_ptr(), 1, 1, desc(m, k), Bt.data_ptr(), 1, 1, desc(k, n), 0.0, Ct.data_ptr(), 1, 1, desc(m, n))
    torch.cuda.synchronize()
    return bool(torch.equal(Ct, Bt @ At))   # small integers: exact


def free_bytes():
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return torch.cuda.mem_get_info()[0]


ok = one(256, 256, 256)                      # loads the kernels, cuBLAS for the check
base = free_bytes()
# (|A| + |B| + |C|) * 8 B: 201, 168, 168, 159, 159, 168 MB -> 1023 MB if every problem stayed cached, <= 369 MB for any two
shapes = [(4096, 4096, 1024), (4096, 2048, 2048), (2048, 4096, 2048), (3072, 4096, 1024), (4096, 3072, 1024), (2048, 2048, 4096)]
for rep in range(2):
    for (m, n, k) in shapes:
        same = one(m, n, k)
        ok = ok and same
        print(rep, m, n, k, "exact", same)
    held = base - free_bytes()
    print("pass", rep, "device memory held by the library: %%.0f MB" %% (held / 2**20))
    ok = ok and held < (700 << 20)
print("RESULT", "OK" if ok else "FAILED")
'''


@pytest.mark.gpu
def test_problem_cache_is_bounded(lib):
    env = dict(os.environ, COSMA_B200_CACHED_PROBLEMS="2")
    out = subprocess.run([sys.executable, "-c", CACHE_SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    text = out.stdout + out.stderr
    assert out.returncode == 0 and "RESULT OK" in text, text[-3000:]


def _run_world_short(world, cases, env, limit=150):
    """tests/test_multiply_gpu.py's multi-GPU worker with extra environment and a short wall-clock limit (a hang must cost seconds)."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_multiply_gpu as tm
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = tm._free_port()
        procs = [ctx.Process(target=tm._worker, args=(r, world, port, cases, q)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(limit)
        hung = [p for p in procs if p.is_alive()]
        for p in hung:
            p.terminate()
        assert not hung, "ranks still running after %d s" % limit
        assert all(p.exitcode == 0 for p in procs)
        res = q.get(timeout=10)
        assert all(res), res
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.gpu
@pytest.mark.parametrize("world,cases", [
    (2, [(1000, 2048, 64, "pn2", "d", 1.0, 0.0, "host"), (1000, 2048, 64, "pn2", "d", 2.0, 1.0, "host"), (512, 1536, 96, "pn2", "z", 1.0, 0.0, "host"),
         (1800, 4096, 64, "pm2", "z", 1.0, 0.5, "host"),                                     # pm2: B is divided more finely than C -> two C pieces per panel
         (1700, 7000, 96, "pk2", "d", 1.0, 1.0, "host")]),                                   # pk2: nothing is gathered -> the plain streamed path
    (4, [(1024, 2048, 512, "pn2,pk2", "d", 1.0, 0.0, "host"), (1024, 2048, 512, "pn2,pk2", "d", 1.0, 1.0, "host")]),
    (8, [(2048, 4096, 2048, "pm2,pn2,pk2", "d", 1.0, 0.0, "host"), (1024, 2048, 1024, "pm2,pn2,pk2", "s", 1.0, 1.0, "host")]),
])
def test_host_panels(lib, world, cases):
    """COSMA_B200_HOST_PANELS=4: cosma_b200_multiply_host as four column panels (A uploaded and gathered once, panels of B up / of C down
    under the GEMMs). Exact against the dense product, twice per case (arenas, streams and events are reused). The cut itself is proven
    on the CPU (test_schedule_cpu.py); seen green on 2 GPUs in profiles/r2b_pytest_gpu_n2.txt."""
    _run_world_short(world, cases, {"COSMA_B200_HOST_PANELS": "4"})
