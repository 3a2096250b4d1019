"""The N > 1 path on CPU: two real processes (torch.distributed, gloo, 127.0.0.1) each compile their own schedule
through the C ABI and execute it with numpy + point-to-point messages; rank 0 gathers C and compares with the dense
product. Exercises ring membership, piece counts and offsets across process boundaries without NCCL."""
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


def _exchange(ring, my_pos, send_to, recv_from):
    """send_to[g] / recv_from[g]: numpy arrays per ring member (None for self). Deadlock-free pairwise exchange."""
    reqs, bufs = [], {}
    for g, q in enumerate(ring):
        if g == my_pos:
            continue
        if recv_from[g] is not None and recv_from[g].size:
            t = torch.empty(recv_from[g].size, dtype=torch.float64)
            bufs[g] = t
            reqs.append(dist.irecv(t, src=q))
        if send_to[g] is not None and send_to[g].size:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(send_to[g])), dst=q))
    for r in reqs:
        r.wait()
    return {g: t.numpy() for g, t in bufs.items()}


def _worker(rank, world, port, m, n, k, steps, alpha, beta, out_q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cosma_b200.distributed import MultiplyPlan, fill_local_from_global, gather_local_to_global
    from schedule_sim import gemm_cpu
    rng = np.random.default_rng(5)
    Ag = rng.integers(0, 10, size=(m, k)).astype(np.float64)
    Bg = rng.integers(0, 10, size=(k, n)).astype(np.float64)
    Cg = rng.integers(0, 10, size=(m, n)).astype(np.float64)
    pl = MultiplyPlan(None, m, n, k, steps, "d", rank=rank, nranks=world, allocate=False)
    arenas = [np.zeros(max(pl.arena_elements[x], 1)) for x in range(3)]
    if not pl.idle:
        for x, (label, full) in enumerate((("A", Ag), ("B", Bg), ("C", Cg))):
            fill_local_from_global(pl, label, arenas[x], full)
        for op in pl.ops():
            if op["kind"] == "gemm":
                gemm_cpu(op, *arenas, alpha, beta)
                continue
            buf, ring, me, piece = arenas[op["matrix"]], op["ring"], op["my_pos"], op["piece"]
            nb, div = len(piece[0]), len(ring)
            if op["kind"] == "allgather":
                mine = buf[op["src_off"]:op["src_off"] + sum(piece[me])].copy()
                got = _exchange(ring, me, [mine if g != me else None for g in range(div)],
                                [np.empty(sum(piece[g])) if g != me else None for g in range(div)])
                got[me] = mine
                pos, off = [0] * div, op["dst_off"]
                for b in range(nb):
                    for g in range(div):
                        c = piece[g][b]
                        buf[off:off + c] = got[g][pos[g]:pos[g] + c]
                        pos[g] += c
                        off += c
            else:
                # slices of my partial result destined to each member, bucket order
                slices = [[] for _ in range(div)]
                off = op["src_off"]
                for b in range(nb):
                    for g in range(div):
                        slices[g].append(buf[off:off + piece[g][b]])
                        off += piece[g][b]
                send = [np.concatenate(s) if s else np.zeros(0) for s in slices]
                got = _exchange(ring, me, [send[g] if g != me else None for g in range(div)],
                                [np.empty(sum(piece[me])) if g != me else None for g in range(div)])
                total = send[me].copy()
                for g in range(div):   # fixed summation order: group index
                    if g != me and g in got:
                        total = total + got[g]
                bt = {0: 0.0, 1: 1.0, 2: beta}[op["beta"]]
                dst = buf[op["dst_off"]:op["dst_off"] + total.size]
                dst[:] = total if bt == 0 else bt * dst + total
    # gather C on rank 0
    mine = torch.from_numpy(arenas[2][:max(pl.initial_elements[2], 1)].copy())
    if rank == 0:
        full = np.zeros((m, n))
        for r in range(pl.P_used):
            if r == 0:
                loc = mine.numpy()
            else:
                cnt = sum((b[1] - b[0] + 1) * (b[3] - b[2] + 1) for b in pl.local_blocks("C", r))
                t = torch.empty(max(cnt, 1), dtype=torch.float64)
                dist.recv(t, src=r)
                loc = t.numpy()
            gather_local_to_global(pl, "C", loc, full, rank=r)
        want = alpha * (Ag @ Bg) + beta * Cg
        out_q.put(bool(np.array_equal(full, want)))
    elif not pl.idle:
        dist.send(mine, dst=0)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("m,n,k,steps,beta", [
    (96, 80, 64, "pk2", 0.0),          # BASELINE configs[0]/[2] strategy at P=2
    (96, 80, 64, "pk2", 1.0),
    (64, 96, 48, "pm2", 1.0),
    (60, 50, 70, "sm2,pn2,sk3", 1.0),  # sequential steps -> several buckets per rank, irregular pieces
])
def test_two_process_gloo(lib, m, n, k, steps, beta):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, m, n, k, steps, 1.0, beta, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
