"""The N > 1 relayout path on CPU: two (and three) real processes (torch.distributed, gloo, 127.0.0.1) each plan
costa::transform for their own rank through the C ABI, pack with the oracle kernel, exchange the per-peer segments
point-to-point, and unpack. Every rank verifies the target blocks it owns against the dense expectation. Exercises
peer byte counts / offsets / piece ordering across real process boundaries without NCCL or a GPU."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dtype, op, out_q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import costa_sim as sim
    from cosma_b200 import costa
    from oracle import oracle
    rng = np.random.default_rng(99)  # same stream on every rank: identical global matrices and layouts
    ok = True
    for trial in range(3):
        m, n = int(rng.integers(8, 120)), int(rng.integers(8, 120))
        sm, sn = (m, n) if op == "N" else (n, m)
        F = sim.DistMatrix(sim.random_split(rng, sm, 4), sim.random_split(rng, sn, 3), rng.integers(0, world, size=(4, 3)), world, dtype,
                           "CR"[trial % 2], pad=1)
        T = sim.DistMatrix(sim.random_split(rng, m, 3), sim.random_split(rng, n, 5), rng.integers(0, world, size=(3, 5)), world, dtype,
                           "RC"[trial % 2], pad=2)
        G, H = sim.random_values(rng, (sm, sn), dtype), sim.random_values(rng, (m, n), dtype)
        F.scatter(G); T.scatter(H)
        alpha, beta = ((1.0, 0.0), (2.0, 1.0), (1.0, -1.0))[trial]
        tp = costa.TransformPlan(None, dtype, [(F.layout(rank), T.layout(rank), op, alpha, beta)], rank=rank, nranks=world)
        pl = tp.export()
        tp.destroy()
        send = np.zeros(max(pl["total_send"], 1), dtype=np.uint8)
        recv = np.zeros(max(pl["total_recv"], 1), dtype=np.uint8)
        sim.run_pieces(oracle, dtype, pl["pack"], [(alpha, beta)], dst_base=send.ctypes.data)
        sim.run_pieces(oracle, dtype, pl["local"], [(alpha, beta)])
        reqs, keep = [], []
        for p in range(world):  # the all-to-all-v (what the NCCL group of send/recv does on the GPU)
            if p == rank:
                continue
            if pl["recv_bytes"][p]:
                t = torch.empty(pl["recv_bytes"][p], dtype=torch.uint8)
                keep.append((p, t))
                reqs.append(dist.irecv(t, src=p))
            if pl["send_bytes"][p]:
                reqs.append(dist.isend(torch.from_numpy(send[pl["send_off"][p]:pl["send_off"][p] + pl["send_bytes"][p]].copy()), dst=p))
        for r in reqs:
            r.wait()
        for p, t in keep:
            recv[pl["recv_off"][p]:pl["recv_off"][p] + pl["recv_bytes"][p]] = t.numpy()
        sim.run_pieces(oracle, dtype, pl["unpack"], [(alpha, beta)], src_base=recv.ctypes.data)
        want = (alpha * sim.apply_op(G, op) + beta * H).astype(sim.NP[dtype])
        got = T.gather()
        for bi in range(len(T.rowsplit) - 1):
            for bj in range(len(T.colsplit) - 1):
                if T.owners[bi, bj] == rank:
                    sl = (slice(T.rowsplit[bi], T.rowsplit[bi + 1]), slice(T.colsplit[bj], T.colsplit[bj + 1]))
                    ok = ok and bool(np.array_equal(got[sl], want[sl]))
        dist.barrier()
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out_q.put(bool(flag.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dtype,op", [(2, "d", "N"), (2, "z", "C"), (3, "d", "T"), (2, "s", "T")])
def test_transform_across_processes(lib, oracle, world, dtype, op):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dtype, op, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=10)
