"""GPU parity of cosma::multiply (compiled schedule + NCCL + sm_100a GEMM) through the C ABI.

Mirrors the reference's distributed test harness (utils/cosma_utils.hpp:226-283): fill local buffers from global
matrices via the Mapper layout, multiply, gather C via the layout, compare with a dense product from the oracle.
Integer-valued inputs -> the comparison is bit-exact (reduction order cannot matter)."""
import os
import time
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


NPDT = {"d": np.float64, "z": np.complex128, "s": np.float32, "c": np.complex64}


def _globals(m, n, k, dtype, seed=11):
    """Integer-valued inputs: exact in FP64 and, below 2^24, in FP32 / 3xTF32 as well -> bit-exact comparisons."""
    rng = np.random.default_rng(seed)
    def r(a, b):
        v = rng.integers(0, 10, size=(a, b)).astype(np.float64)
        if dtype in "zc":
            v = v + 1j * rng.integers(0, 10, size=(a, b))
        return v.astype(NPDT[dtype])
    return r(m, k), r(k, n), r(m, n)


def _dense_oracle(oracle, Ag, Bg, Cg, alpha, beta):
    m, k = Ag.shape
    n = Bg.shape[1]
    if Ag.dtype in (np.float32, np.complex64):  # exact integers: form the expected result in double, then narrow
        wide = np.complex128 if Ag.dtype == np.complex64 else np.float64
        return (alpha * (Ag.astype(wide) @ Bg.astype(wide)) + beta * Cg.astype(wide)).astype(Ag.dtype)
    C = np.ascontiguousarray(Cg.T).reshape(-1).copy()
    out = oracle.gemm("N", "N", m, n, k, alpha, np.ascontiguousarray(Ag.T).reshape(-1), m, np.ascontiguousarray(Bg.T).reshape(-1), k,
                      beta, C, m)
    return out.reshape(n, m).T


@pytest.mark.parametrize("dtype", ["d", "z", "s", "c"])
@pytest.mark.parametrize("m,n,k,steps,alpha,beta", [
    (256, 192, 320, "", 1.0, 0.0),
    (300, 200, 100, "sm2,sn2,sk2", 1.0, 1.0),       # sequential steps: bucket offsets, beta = 1 from the 2nd k chunk on
    (130, 141, 152, "sk3,sm2", 2.0, -1.0),          # irregular splits, odd leading dimensions (generic GEMM path too)
    (64, 64, 64, "sn4", 1.0, 0.0),
])
def test_single_gpu_schedules(oracle, dtype, m, n, k, steps, alpha, beta):
    from cosma_b200.distributed import MultiplyPlan, fill_local_from_global, gather_local_to_global
    Ag, Bg, Cg = _globals(m, n, k, dtype)
    pl = MultiplyPlan(None, m, n, k, steps, dtype, rank=0, nranks=1)
    for label, mat, full in (("A", pl.A, Ag), ("B", pl.B, Bg), ("C", pl.C, Cg)):
        host = np.zeros(mat.initial, dtype=full.dtype)
        fill_local_from_global(pl, label, host, full)
        mat.local.copy_(torch.from_numpy(host))
    if beta == 0.0:
        pl.C.local.fill_(float("nan"))
    pl.multiply(alpha, beta)
    torch.cuda.synchronize()
    got = np.zeros((m, n), dtype=Ag.dtype)
    gather_local_to_global(pl, "C", pl.C.local.cpu().numpy(), got)
    want = _dense_oracle(oracle, Ag, Bg, Cg, alpha, beta)
    assert np.array_equal(got, want)
    pl.destroy()


def test_single_gpu_host_entry_point(oracle):
    from cosma_b200.distributed import MultiplyPlan, fill_local_from_global, gather_local_to_global
    m, n, k = 200, 150, 100
    Ag, Bg, Cg = _globals(m, n, k, "d")
    pl = MultiplyPlan(None, m, n, k, "sk2", "d", rank=0, nranks=1)
    hs = []
    for label, full in (("A", Ag), ("B", Bg), ("C", Cg)):
        h = torch.zeros(pl.initial_elements["ABC".index(label)], dtype=torch.float64).pin_memory()
        fill_local_from_global(pl, label, h.numpy(), full)
        hs.append(h)
    pl.multiply_host(hs[0], hs[1], hs[2], 1.0, 1.0)
    torch.cuda.synchronize()
    got = np.zeros((m, n))
    gather_local_to_global(pl, "C", hs[2].numpy(), got)
    assert np.array_equal(got, _dense_oracle(oracle, Ag, Bg, Cg, 1.0, 1.0))
    pl.destroy()


@pytest.mark.parametrize("dtype", ["d", "z", "s", "c"])
@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_single_gpu_host_entry_point_streamed(oracle, dtype, beta):
    """P = 1, one GEMM: A in row chunks under a wide first panel, B / C in column panels (host_gemm.cu). Sizes span
    several row chunks (768) and panels (first 3072 / 1536 / 5632 / 3072 for d / z / s / c, then 2048, thin tail)."""
    from cosma_b200.distributed import MultiplyPlan
    m, k = 1700, 48
    n = {"d": 3072 + 2048 + 1700, "z": 1536 + 2048 + 1100, "s": 5632 + 2048 + 300, "c": 3072 + 2048 + 2048 + 1600}[dtype]
    Ag, Bg, Cg = _globals(m, n, k, dtype)
    pl = MultiplyPlan(None, m, n, k, "", dtype, rank=0, nranks=1)
    tdt = {"d": torch.float64, "z": torch.complex128, "s": torch.float32, "c": torch.complex64}[dtype]
    hs = [torch.from_numpy(np.ascontiguousarray(x.T).reshape(-1).copy()).to(tdt).pin_memory() for x in (Ag, Bg, Cg)]
    if beta == 0.0:
        hs[2].fill_(float("nan"))
    for _ in range(2):  # twice: buffers and events are reused
        if beta != 0.0:
            hs[2].copy_(torch.from_numpy(np.ascontiguousarray(Cg.T).reshape(-1)))
        pl.multiply_host(hs[0], hs[1], hs[2], 1.0, beta)
        torch.cuda.synchronize()
    got = hs[2].numpy().reshape(n, m).T
    wide = np.complex128 if dtype in "zc" else np.float64
    want = (Ag.astype(wide) @ Bg.astype(wide) + beta * Cg.astype(wide)).astype(Ag.dtype)
    assert np.array_equal(got, want)
    if dtype == "d":  # and against the oracle proper
        assert np.array_equal(got, _dense_oracle(oracle, Ag, Bg, Cg, 1.0, beta))
    pl.destroy()


# ---- multi-GPU: one process per GPU, NCCL ---------------------------------------------------------------------------

def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, cases, q):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from cosma_b200.distributed import init_comm, MultiplyPlan, fill_local_from_global, gather_local_to_global
    comm = init_comm()
    results = []
    for case in cases:
        (m, n, k, steps, dtype, alpha, beta), tags = case[:7], set(case[7:])
        via_host = "host" in tags
        Ag, Bg, Cg = _globals(m, n, k, dtype)
        # "overlap": the communication / computation overlap forced on whatever its estimated gain (the decision is taken at plan
        # creation); "serial": switched off. Untagged cases take the library's own decision.
        os.environ.pop("COSMA_OVERLAP_COMM_AND_COMP", None)
        if "overlap" in tags:
            os.environ["COSMA_OVERLAP_COMM_AND_COMP"] = "FORCE"
        elif "serial" in tags:
            os.environ["COSMA_OVERLAP_COMM_AND_COMP"] = "OFF"
        pl = MultiplyPlan(comm, m, n, k, steps, dtype)
        os.environ.pop("COSMA_OVERLAP_COMM_AND_COMP", None)
        if "overlap" in tags and not pl.idle:
            assert pl.overlap()["enabled"], pl.overlap()["why"]
        if "serial" in tags:
            assert not pl.overlap()["enabled"]
        if not pl.idle:
            for label, mat, full in (("A", pl.A, Ag), ("B", pl.B, Bg), ("C", pl.C, Cg)):
                host = np.zeros(mat.initial, dtype=full.dtype)
                fill_local_from_global(pl, label, host, full)
                mat.local.copy_(torch.from_numpy(host))
            if beta == 0.0:
                pl.C.local.fill_(float("nan"))
        for _ in range(2):           # run twice: plans and ring communicators are reusable
            if not pl.idle and _ == 1:
                host = np.zeros(pl.C.initial, dtype=Cg.dtype); fill_local_from_global(pl, "C", host, Cg)
                pl.C.local.copy_(torch.from_numpy(host))
                if beta == 0.0:
                    pl.C.local.fill_(float("nan"))
            if via_host and not pl.idle:
                # host-pointer entry point: pinned local matrices in, local C out (un-gathered operands are streamed)
                hs = [t.local.cpu().pin_memory() for t in (pl.A, pl.B, pl.C)]
                pl.multiply_host(hs[0], hs[1], hs[2], alpha, beta)
                torch.cuda.synchronize()
                pl.C.local.copy_(hs[2])
            else:
                pl.multiply(alpha, beta)
        torch.cuda.synchronize()
        # gather the local C buffers on rank 0 through torch.distributed (test plumbing only)
        tdt = {"d": torch.float64, "z": torch.complex128, "s": torch.float32, "c": torch.complex64}[dtype]
        sizes = [sum((b[1] - b[0] + 1) * (b[3] - b[2] + 1) for b in pl.local_blocks("C", r)) for r in range(world)]
        mx = max(max(sizes), 1)
        mine = torch.zeros(mx, dtype=tdt, device="cuda")
        if not pl.idle:
            mine[:pl.C.initial] = pl.C.local
        real = torch.view_as_real(mine).reshape(-1) if dtype in "zc" else mine
        allb = [torch.empty_like(real) for _ in range(world)]
        dist.all_gather(allb, real)
        if rank == 0:
            got = np.zeros((m, n), dtype=Cg.dtype)
            for r in range(pl.P_used):
                loc = allb[r].cpu().numpy()
                loc = loc.view(NPDT[dtype]) if dtype in "zc" else loc
                gather_local_to_global(pl, "C", loc, got, rank=r)
            wide = np.complex128 if dtype in "zc" else np.float64
            results.append((got, (alpha * (Ag.astype(wide) @ Bg.astype(wide)) + beta * Cg.astype(wide)).astype(Cg.dtype)))
        pl.destroy()
    if rank == 0:
        q.put([bool(np.array_equal(g, w)) for g, w in results])
    dist.barrier()
    comm.destroy()
    dist.destroy_process_group()


def _run_world(world, cases):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    # one wall-clock limit for the whole world: a rank stuck in a collective must cost minutes, not the GPU call
    deadline = time.time() + 300
    for p in procs:
        p.join(max(1.0, deadline - time.time()))
    hung = [p for p in procs if p.is_alive()]
    for p in hung:
        p.terminate()
    assert not hung, "ranks still running after the time limit"
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = q.get(timeout=10)
    assert all(res), res


def test_two_gpus(lib):
    _run_world(2, [
        (512, 512, 512, "", "d", 1.0, 0.0),                 # -> automatic strategy (pk2 shape of BASELINE configs[0]/[2])
        (2000, 2000, 2000, "pk2", "d", 1.0, 0.0),           # BASELINE configs[0]
        (640, 512, 384, "pm2", "d", 1.0, 1.0),
        (300, 260, 220, "sm2,pn2,sk3", "d", 2.0, 1.0),      # several buckets per rank -> exact-count grouped send/recv
        (257, 129, 511, "pk2", "d", 1.0, 1.0),              # irregular k split, beta != 0 -> staged reduce + axpby
        (256, 256, 256, "pk2", "z", 1.0 - 0.5j, 0.5j),
        (512, 384, 256, "pk2", "s", 1.0, 1.0),              # single precision: 3xTF32 tcgen05 base case, float NCCL reduce
        (300, 260, 220, "sm2,pn2,sk3", "c", 2.0, 1.0),
        # cosma_b200_multiply_host: A and B streamed under the GEMM (pk2); only B (pn2) / only A (pm2) with C streamed out
        (1700, 7000, 96, "pk2", "d", 1.0, 0.0, "host"),
        (1700, 7000, 96, "pk2", "d", 1.0, 1.0, "host"),
        (1000, 9000, 64, "pn2", "d", 1.0, 1.0, "host"),
        (1800, 4000, 64, "pm2", "z", 1.0, 0.0, "host"),
        (300, 260, 220, "sm2,pn2,sk3", "d", 2.0, 1.0, "host"),  # several GEMMs: plain up-front copies
        # communication / computation overlap (COSMA_OVERLAP_COMM_AND_COMP): each split kind on its own
        (1024, 1024, 1024, "pk2", "d", 1.0, 0.0, "overlap"),    # peer's half of C first, exchanged while the own half is computed
        (1024, 1024, 1024, "pk2", "d", 2.0, -1.0, "overlap"),   # ... with the staged beta term
        (1024, 1024, 1024, "pm2", "d", 1.0, 1.0, "overlap"),    # own column block of B first
        (1024, 1024, 1024, "pn2", "z", 1.0 - 0.5j, 0.5j, "overlap"),  # own k block of A first, the other accumulated
        (2048, 1024, 512, "pk2", "s", 1.0, 1.0, "overlap"),
        (1024, 1024, 1024, "pk2", "d", 1.0, 0.0, "serial"),
    ])


def test_four_gpus(lib):
    _run_world(4, [
        (1024, 1024, 1024, "pn2,pk2", "d", 1.0, 0.0),       # BASELINE configs[2] strategy at P=4
        (100, 100, 100, "pm2,pk2", "d", 1.0, 1.0),          # reference tests/multiply.cpp case
        (20, 30, 25, "sm2,sn2,pk2,pm2", "d", 1.0, 1.0),
        (400, 400, 400, "", "z", 1.0, 0.0),
        (512, 512, 512, "pn2,pk2", "s", 1.0, 0.0),
        (1024, 1024, 1024, "pn2,pk2", "d", 1.0, 0.0, "overlap"),
        (1024, 2048, 1024, "pm2,pk2", "d", 2.0, 1.0, "overlap"),
        (1024, 1024, 1024, "pm2,pn2", "z", 1.0, 1.0, "overlap"),
        (1024, 1024, 1024, "pk2,pk2", "c", 1.0, 0.0, "overlap"),   # inner reduce overlapped, outer one serial
    ])


def test_eight_gpus(lib):
    _run_world(8, [
        (2048, 2048, 2048, "pm2,pn2,pk2", "d", 1.0, 0.0),   # BASELINE configs[2] strategy at P=8
        (512, 512, 16384, "pk8", "d", 1.0, 0.0),            # BASELINE configs[3] strategy
        (100, 100, 100, "sm2,pn2,sk2,pm2,sn2,pk2", "d", 1.0, 1.0),  # tests/scalar_matmul.cpp
        (200, 200, 200, "sk3,sm3,sn3,pk2,pn2,pm2", "d", 1.0, 1.0),  # tests/multiply.cpp
        (512, 32, 736, "pk2,pm2,pk2", "d", 1.0, 1.0),               # tests/multiply.cpp (nested k reductions)
        (1000, 1000, 1000, "pm2,pn2,pk2", "z", 1.0, 1.0),
        (1024, 1024, 1024, "pm2,pn2,pk2", "s", 1.0, 0.0),
        (100, 100, 100, "sm2,pn2,sk2,pm2,sn2,pk2", "c", 1.0, 1.0),  # tests/scalar_matmul.cpp, complex<float>
        (2048, 2048, 2048, "pm2,pn2,pk2", "d", 1.0, 0.0, "overlap"),   # BASELINE configs[2] strategy, both allgathers and the reduce overlapped
        (2048, 2048, 2048, "pm2,pn2,pk2", "d", 2.0, 1.0, "overlap"),
        (2048, 1024, 2048, "pn2,pm2,pk2", "z", 1.0, 0.5, "overlap"),
        (2048, 2048, 1024, "pk2,pm2,pn2", "s", 1.0, 0.0, "overlap"),   # k split first: the reduce still follows the GEMM directly
    ])
