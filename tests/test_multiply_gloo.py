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


# ---- the overlapped schedules (include/cosma/overlap.hpp) across real processes ---------------------------------------------

def _worker_overlapped(rank, world, port, m, n, k, steps, alpha, beta, zero_sm, out_q):
    """Every process interprets ITS micro-op program (MultiplyPlan.overlap(): GEMM panels, allgathers, the exchange of the partial C
    halves, accumulations) in program order -- a valid serialisation of its two streams -- with gloo point-to-point messages where
    the device executor pushes through copy engines / NCCL."""
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["COSMA_OVERLAP_COMM_AND_COMP"] = "FORCE"
    os.environ["COSMA_B200_OVERLAP_GRANULE"] = "8"
    if zero_sm:
        os.environ["COSMA_B200_OVERLAP_ZERO_SM"] = "ON"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cosma_b200.distributed import MultiplyPlan, fill_local_from_global, gather_local_to_global
    from schedule_sim import micro_gemm_cpu
    rng = np.random.default_rng(9)
    Ag = rng.integers(0, 10, size=(m, k)).astype(np.float64)
    Bg = rng.integers(0, 10, size=(k, n)).astype(np.float64)
    Cg = rng.integers(0, 10, size=(m, n)).astype(np.float64)
    pl = MultiplyPlan(None, m, n, k, steps, "d", rank=rank, nranks=world, allocate=False)
    ov = pl.overlap()
    assert ov["enabled"], ov["why"]
    sched = pl.ops()
    arenas = [np.full(max(pl.arena_elements[x], 1), np.nan) for x in range(3)]   # poisoned: nothing is read before it arrives
    for x, (label, full) in enumerate((("A", Ag), ("B", Bg), ("C", Cg if beta != 0 else np.full_like(Cg, np.nan)))):
        arenas[x][:pl.initial_elements[x]] = 0.0
        fill_local_from_global(pl, label, arenas[x], full)
    bt = lambda mode: {0: 0.0, 1: 1.0, 2: beta}[mode]
    for o in ov["ops"]:
        if o["kind"] == "gemm":
            micro_gemm_cpu(o, *arenas, alpha, beta)
        elif o["kind"] == "accumulate":
            b = bt(o["beta"])
            if not (o["beta_term"] and b == 0):
                C = arenas[2]
                C[o["dst_off"]:o["dst_off"] + o["count"]] = b * C[o["dst_off"]:o["dst_off"] + o["count"]] + C[o["add_off"]:o["add_off"] + o["count"]]
        elif o["kind"] == "allgather":
            op = sched[o["op"]]
            assert op["regular"] and len(op["ring"]) == 2
            buf, me, cnt = arenas[op["matrix"]], op["my_pos"], op["piece"][0][0]
            mine = buf[op["src_off"]:op["src_off"] + cnt].copy()
            got = _exchange(op["ring"], me, [mine if g != me else None for g in range(2)], [np.empty(cnt) if g != me else None for g in range(2)])
            buf[op["dst_off"] + me * cnt:op["dst_off"] + (me + 1) * cnt] = mine
            buf[op["dst_off"] + (1 - me) * cnt:op["dst_off"] + (2 - me) * cnt] = got[1 - me]
        elif o["kind"] == "exchange":
            ring = next(s["ring"] for s in sched if s["kind"] != "gemm" and s["ring_index"] == o["ring_index"])
            me = 1 - o["peer"]
            C = arenas[2]
            send = C[o["send_off"]:o["send_off"] + o["count"]].copy()
            got = _exchange(ring, me, [send if g != me else None for g in range(2)], [np.empty(o["count"]) if g != me else None for g in range(2)])
            at = o["recv_off_zero"] if bt(o["beta"]) == 0 else o["recv_off"]
            C[at:at + o["count"]] = got[o["peer"]]
        else:
            raise AssertionError("serial collectives are not part of these cases: %s" % o)
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
    else:
        dist.send(mine, dst=0)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,m,n,k,steps,beta,zero_sm", [
    (2, 64, 64, 64, "pk2", 2.0, False), (2, 64, 64, 64, "pk2", 0.0, True), (2, 64, 96, 64, "pm2", 1.0, True), (2, 96, 64, 64, "pn2", 0.0, False),
    (4, 128, 128, 128, "pn2,pk2", 0.0, True), (4, 128, 128, 128, "pm2,pn2", 1.0, False),
])
def test_overlapped_program_across_processes(lib, world, m, n, k, steps, beta, zero_sm):
    """zero_sm: the panel shapes of the copy-engine transport (no narrow launch), else those of the NCCL transport."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_overlapped, args=(r, world, port, m, n, k, steps, 2.0, beta, zero_sm, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
