"""Host-side mirror of cosma::multiply for one-process-per-GPU jobs (torch.distributed for rendezvous only).

  comm = init_comm()                              # ncclUniqueId broadcast over torch.distributed -> ncclCommInitRank
  plan = MultiplyPlan(comm, m, n, k, steps="", dtype="d")
  plan.A.local[...] = ...                         # the rank's local matrix in the reference's layout
  plan.multiply(alpha, beta)                      # C = alpha*A*B + beta*C, asynchronous on the current stream

Mirrors: CosmaMatrix<T> A('A', strategy, rank) / matrix_pointer() / matrix_size() and
multiply(A, B, C, strategy, comm, alpha, beta) (reference src/cosma/matrix.hpp:26-213, multiply.hpp:47-54)."""
import ctypes

from . import _lib

_MAT = {"A": 0, "B": 1, "C": 2}


def _declare(lib):
    if getattr(lib, "_dist_declared", False):
        return
    vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.cosma_b200_comm_create.argtypes = [ci, ci, ctypes.c_char_p, ctypes.POINTER(vp)]
    lib.cosma_b200_comm_destroy.argtypes = [vp]
    lib.cosma_b200_plan_create.argtypes = [vp, ci, ci, ci, ci, ci, ctypes.c_char_p, ctypes.c_char, ctypes.POINTER(vp)]
    lib.cosma_b200_plan_destroy.argtypes = [vp]
    lib.cosma_b200_plan_arena_elements.argtypes = [vp, ci]
    lib.cosma_b200_plan_arena_elements.restype = i64
    lib.cosma_b200_plan_initial_elements.argtypes = [vp, ci]
    lib.cosma_b200_plan_initial_elements.restype = i64
    lib.cosma_b200_plan_strategy.argtypes = [vp, ctypes.c_char_p, ci, ctypes.POINTER(ci)]
    lib.cosma_b200_plan_gemm_flops.argtypes = [vp]
    lib.cosma_b200_plan_gemm_flops.restype = ctypes.c_double
    lib.cosma_b200_plan_export.argtypes = [vp, ctypes.POINTER(i64), i64, ctypes.POINTER(i64)]
    lib.cosma_b200_plan_local_blocks.argtypes = [vp, ci, ci, ctypes.POINTER(ci), ci, ctypes.POINTER(ci)]
    lib.cosma_b200_multiply.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), vp, vp, vp, vp]
    lib.cosma_b200_multiply_host.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), vp, vp, vp, vp]
    lib.cosma_b200_plan_last_launches.argtypes = [vp]
    lib.cosma_b200_plan_time_gemms.argtypes = [vp, ci]
    lib.cosma_b200_plan_gemm_times.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ci, ctypes.POINTER(ci)]
    lib.cosma_b200_plan_bind_arenas.argtypes = [vp, vp, vp, vp, ctypes.POINTER(ci)]
    lib.cosma_b200_plan_overlap_export.argtypes = [vp, ctypes.POINTER(i64), i64, ctypes.POINTER(i64), ctypes.POINTER(ci), ctypes.c_char_p, ci,
                                                   ctypes.POINTER(ctypes.c_double)]
    lib._dist_declared = True


class Comm:
    """NCCL communicator of the whole job (the reference's MPI_Comm argument)."""

    def __init__(self, handle, rank, size):
        self.handle, self.rank, self.size = handle, rank, size

    def destroy(self):
        if self.handle:
            _lib.load().cosma_b200_comm_destroy(self.handle)
            self.handle = None


def init_comm(device=None):
    """Collective over torch.distributed's default group. Single-process jobs get a trivial communicator."""
    import torch
    import torch.distributed as dist
    lib = _lib.load()
    _declare(lib)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        handle = ctypes.c_void_p()
        _lib.check(lib.cosma_b200_comm_create(0, 1, None, ctypes.byref(handle)), "cosma_b200_comm_create")
        return Comm(handle, 0, 1)
    rank, size = dist.get_rank(), dist.get_world_size()
    uid = (ctypes.c_uint8 * 128)()
    if rank == 0:
        _lib.check(lib.cosma_b200_nccl_unique_id(uid), "cosma_b200_nccl_unique_id")
    on_gpu = dist.get_backend() == "nccl"
    t = torch.tensor(list(uid), dtype=torch.uint8, device=(device or torch.device("cuda", torch.cuda.current_device())) if on_gpu else "cpu")
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    handle = ctypes.c_void_p()
    _lib.check(lib.cosma_b200_comm_create(rank, size, raw, ctypes.byref(handle)), "cosma_b200_comm_create")
    return Comm(handle, rank, size)


def parse_plan(flat):
    """Decodes cosma_b200_plan_export (see csrc/host/schedule.cpp) into a list of dicts."""
    ops, i = [], 0
    while i < len(flat):
        kind = flat[i]; i += 1
        if kind == 0:
            a, b, c, m, n, k, beta = flat[i:i + 7]; i += 7
            ops.append({"kind": "gemm", "a_off": a, "b_off": b, "c_off": c, "m": m, "n": n, "k": k, "beta": beta})
            continue
        matrix, step, ring_index, my_pos, src, dst = flat[i:i + 6]; i += 6
        op = {"kind": "allgather" if kind == 1 else "reduce", "matrix": matrix, "step": step, "ring_index": ring_index,
              "my_pos": my_pos, "src_off": src, "dst_off": dst}
        if kind == 2:
            op["tmp_off"], op["beta"] = flat[i], flat[i + 1]; i += 2
        div, nb, regular = flat[i:i + 3]; i += 3
        op["regular"] = bool(regular)
        op["ring"] = list(flat[i:i + div]); i += div
        op["piece"] = [list(flat[i + g * nb:i + (g + 1) * nb]) for g in range(div)]; i += div * nb
        ops.append(op)
    return ops


def parse_overlap(flat):
    """Decodes cosma_b200_plan_overlap_export (include/cosma/overlap.hpp, OverlapProgram::serialize) into a list of dicts."""
    ops, i = [], 0
    names = ("gemm", "allgather", "exchange", "accumulate", "serial")
    while i < len(flat):
        kind, stream, nw = flat[i:i + 3]; i += 3
        op = {"kind": names[kind], "stream": stream, "wait": list(flat[i:i + nw])}; i += nw
        if kind == 0:
            keys = ("a_off", "b_off", "c_off", "lda", "ldb", "ldc", "m", "n", "k", "beta", "narrow")
        elif kind in (1, 4):
            keys = ("op",)
        elif kind == 2:
            keys = ("ring_index", "peer", "send_off", "recv_off", "recv_off_zero", "count", "beta")
        else:
            keys = ("dst_off", "add_off", "count", "beta", "beta_term")
        op.update(zip(keys, flat[i:i + len(keys)])); i += len(keys)
        ops.append(op)
    return ops


class LocalMatrix:
    """One matrix of a plan on this rank: device arena + the view of the rank's local data (CosmaMatrix analogue)."""

    def __init__(self, label, arena, initial):
        self.label, self.arena, self.initial = label, arena, initial

    @property
    def local(self):  # matrix_pointer() / matrix_size()
        return self.arena[:self.initial]


class MultiplyPlan:
    def __init__(self, comm, m, n, k, steps="", dtype="d", rank=None, nranks=None, device=None, allocate=True):
        lib = _lib.load()
        _declare(lib)
        self.lib, self.comm, self.m, self.n, self.k, self.dtype = lib, comm, m, n, k, dtype
        self.rank = comm.rank if comm is not None and rank is None else (rank or 0)
        self.nranks = comm.size if comm is not None and nranks is None else (nranks or 1)
        h = ctypes.c_void_p()
        ch = comm.handle if comm is not None else None
        _lib.check(lib.cosma_b200_plan_create(ch, self.rank, self.nranks, m, n, k, steps.encode(), ctypes.c_char(dtype.encode()),
                                              ctypes.byref(h)), "cosma_b200_plan_create")
        self.handle = h
        buf = ctypes.create_string_buffer(4096)
        P = ctypes.c_int(0)
        _lib.check(lib.cosma_b200_plan_strategy(h, buf, 4096, ctypes.byref(P)), "cosma_b200_plan_strategy")
        self.strategy, self.P_used = buf.value.decode(), P.value
        self.idle = self.rank >= self.P_used
        self.arena_elements = [lib.cosma_b200_plan_arena_elements(h, x) for x in range(3)]
        self.initial_elements = [lib.cosma_b200_plan_initial_elements(h, x) for x in range(3)]
        self.gemm_flops = lib.cosma_b200_plan_gemm_flops(h)
        self.A = self.B = self.C = None
        if allocate:
            import torch
            tdt = {"d": torch.float64, "z": torch.complex128, "s": torch.float32, "c": torch.complex64}[dtype]
            dev = device or torch.device("cuda", torch.cuda.current_device())
            mats = []
            for x, label in enumerate("ABC"):
                arena = torch.zeros(max(self.arena_elements[x], 1), dtype=tdt, device=dev)
                mats.append(LocalMatrix(label, arena, self.initial_elements[x]))
            self.A, self.B, self.C = mats
            # the arenas are fixed for the plan's life: overlapped transfers can go through copy engines into the ring mates' arenas
            # (collective over the communicator; nothing changes for plans that are not overlapped)
            self.peer_copy = False
            if comm is not None and self.nranks > 1:
                act = ctypes.c_int(0)
                _lib.check(lib.cosma_b200_plan_bind_arenas(h, ctypes.c_void_p(self.A.arena.data_ptr()), ctypes.c_void_p(self.B.arena.data_ptr()),
                                                           ctypes.c_void_p(self.C.arena.data_ptr()), ctypes.byref(act)), "cosma_b200_plan_bind_arenas")
                self.peer_copy = bool(act.value)

    def ops(self):
        n = ctypes.c_int64(0)
        self.lib.cosma_b200_plan_export(self.handle, None, 0, ctypes.byref(n))
        buf = (ctypes.c_int64 * max(n.value, 1))()
        self.lib.cosma_b200_plan_export(self.handle, buf, n.value, ctypes.byref(n))
        return parse_plan(list(buf[:n.value]))

    def overlap(self):
        """-> {"enabled", "why", "ops": micro-ops, "est_ms": (serial, overlapped, communication)}: how the plan overlaps its
        collectives with the local GEMM (COSMA_OVERLAP_COMM_AND_COMP), see include/cosma/overlap.hpp."""
        n, en = ctypes.c_int64(0), ctypes.c_int(0)
        why = ctypes.create_string_buffer(512)
        est = (ctypes.c_double * 3)()
        _lib.check(self.lib.cosma_b200_plan_overlap_export(self.handle, None, 0, ctypes.byref(n), ctypes.byref(en), why, 512, est), "cosma_b200_plan_overlap_export")
        buf = (ctypes.c_int64 * max(n.value, 1))()
        _lib.check(self.lib.cosma_b200_plan_overlap_export(self.handle, buf, n.value, ctypes.byref(n), ctypes.byref(en), why, 512, est), "cosma_b200_plan_overlap_export")
        return {"enabled": bool(en.value), "why": why.value.decode(), "ops": parse_overlap(list(buf[:n.value])) if en.value else [], "est_ms": tuple(est)}

    def local_blocks(self, label, rank=None):
        """Blocks (row_first,row_last,col_first,col_last) of `rank` (default: this rank) in local-buffer order."""
        rank = self.rank if rank is None else rank
        if rank >= self.P_used:
            return []
        n = ctypes.c_int(0)
        self.lib.cosma_b200_plan_local_blocks(self.handle, _MAT[label], rank, None, 0, ctypes.byref(n))
        out = (ctypes.c_int * max(4 * n.value, 1))()
        self.lib.cosma_b200_plan_local_blocks(self.handle, _MAT[label], rank, out, 4 * n.value, ctypes.byref(n))
        return [tuple(out[4 * i:4 * i + 4]) for i in range(n.value)]

    def multiply(self, alpha=1.0, beta=0.0, stream=None):
        import torch
        if self.dtype in "ds":
            al = (ctypes.c_double * 1)(float(alpha)); be = (ctypes.c_double * 1)(float(beta))
        else:
            al = (ctypes.c_double * 2)(complex(alpha).real, complex(alpha).imag)
            be = (ctypes.c_double * 2)(complex(beta).real, complex(beta).imag)
        s = stream if stream is not None else torch.cuda.current_stream()
        st = self.lib.cosma_b200_multiply(self.handle, al, be, ctypes.c_void_p(self.A.arena.data_ptr()),
                                          ctypes.c_void_p(self.B.arena.data_ptr()), ctypes.c_void_p(self.C.arena.data_ptr()),
                                          ctypes.c_void_p(s.cuda_stream))
        _lib.check(st, "cosma_b200_multiply")
        return self.lib.cosma_b200_plan_last_launches(self.handle)

    def multiply_host(self, hA, hB, hC, alpha=1.0, beta=0.0, stream=None):
        """hA, hB, hC: pinned host tensors holding the rank's local matrices (reference layout)."""
        import torch
        if self.dtype in "ds":
            al = (ctypes.c_double * 1)(float(alpha)); be = (ctypes.c_double * 1)(float(beta))
        else:
            al = (ctypes.c_double * 2)(complex(alpha).real, complex(alpha).imag)
            be = (ctypes.c_double * 2)(complex(beta).real, complex(beta).imag)
        s = stream if stream is not None else torch.cuda.current_stream()
        st = self.lib.cosma_b200_multiply_host(self.handle, al, be, ctypes.c_void_p(hA.data_ptr()), ctypes.c_void_p(hB.data_ptr()),
                                               ctypes.c_void_p(hC.data_ptr()), ctypes.c_void_p(s.cuda_stream))
        _lib.check(st, "cosma_b200_multiply_host")

    def time_gemms(self, enable=True):
        self.lib.cosma_b200_plan_time_gemms(self.handle, 1 if enable else 0)

    def gemm_times_ms(self):
        n = ctypes.c_int(0)
        out = (ctypes.c_float * 4096)()
        st = self.lib.cosma_b200_plan_gemm_times(self.handle, out, 4096, ctypes.byref(n))
        _lib.check(st, "cosma_b200_plan_gemm_times")
        return list(out[:n.value])

    def op_times(self):
        """[(kind, ms, wire_bytes)] of the last synchronised run with timing on; kind in 'gemm' | 'allgather' | 'reduce'."""
        n = ctypes.c_int(0)
        cap = 65536
        kinds, ms, wb = (ctypes.c_int * cap)(), (ctypes.c_float * cap)(), (ctypes.c_int64 * cap)()
        self.lib.cosma_b200_plan_op_times.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float),
                                                      ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        _lib.check(self.lib.cosma_b200_plan_op_times(self.handle, kinds, ms, wb, cap, ctypes.byref(n)), "cosma_b200_plan_op_times")
        names = ("gemm", "allgather", "reduce", "accumulate")
        return [(names[kinds[i]], ms[i], wb[i]) for i in range(min(n.value, cap))]

    def destroy(self):
        if self.handle:
            self.lib.cosma_b200_plan_destroy(self.handle)
            self.handle = None


def fill_local_from_global(plan, label, local, full):
    """Scatter: copies this rank's blocks of the (rows x cols, torch or numpy, indexable [i, j]) global matrix into
    the local buffer, block after block, column-major inside a block -- the reference's layout."""
    pos = 0
    for (r0, r1, c0, c1) in plan.local_blocks(label):
        blk = full[r0:r1 + 1, c0:c1 + 1]
        cnt = (r1 - r0 + 1) * (c1 - c0 + 1)
        local[pos:pos + cnt] = blk.T.reshape(-1)
        pos += cnt
    return pos


def gather_local_to_global(plan, label, local, full, rank=None):
    pos = 0
    for (r0, r1, c0, c1) in plan.local_blocks(label, rank):
        nr, nc = r1 - r0 + 1, c1 - c0 + 1
        full[r0:r1 + 1, c0:c1 + 1] = local[pos:pos + nr * nc].reshape(nc, nr).T
        pos += nr * nc
    return pos


def _formula(which, rows, cols, cplx):
    """Integer-valued synthetic operands defined by GLOBAL coordinates, so that any rank can evaluate any element of A (which = 0) or
    B (which = 1) without communication: values in 0..9 (Tiled-MM's exact-test convention, libs/Tiled-MM/tests/test-multiply.cpp:60-68).
    rows, cols: int64 tensors that broadcast against each other. -> (re, im) int64 tensors (im = None for real types)."""
    def f(salt):
        h = (rows * 1664525 + cols * 1013904223 + (rows ^ cols) * 69069 + salt * 7919) & 0x7FFFFFFF
        return (h >> 7) % 10
    return f(which), (f(which + 2) if cplx else None)


class MultiplyJob:
    """bench.py's workload: one cosma::multiply of (m, n, k) over `world` GPUs (world = 1: a single local GEMM) with the automatic
    strategy, synthetic U[0,10) local matrices (the reference miniapp's fill, miniapp/cosma_miniapp.cpp:21-25,64-70)."""

    def __init__(self, m, n, k, world, rank, device, steps="", dtype="d", comm=None):
        import torch
        self.own_comm = comm is None
        self.comm = comm if comm is not None else init_comm(device)
        self.dtype = dtype
        self.tdt = {"d": torch.float64, "z": torch.complex128, "s": torch.float32, "c": torch.complex64}[dtype]
        self.cplx = dtype in "zc"
        self.plan = MultiplyPlan(self.comm, m, n, k, steps, dtype, device=device)
        self.strategy_string = self.plan.strategy
        gen = torch.Generator(device=device)
        gen.manual_seed(rank)
        rdt = torch.float64 if dtype in "dz" else torch.float32
        for mat in (self.plan.A, self.plan.B):
            if mat.initial:
                if self.cplx:
                    mat.local.copy_(torch.view_as_complex(torch.rand(mat.initial, 2, device=device, dtype=rdt, generator=gen) * 10))
                else:
                    mat.local.copy_(torch.rand(mat.initial, device=device, dtype=rdt, generator=gen) * 10)
        self.plan.C.local.fill_(float("nan"))
        self.plan.time_gemms(True)
        self.flop_factor = 8.0 if self.cplx else 2.0
        self.device, self.rank, self.world = device, rank, world

    def run(self):
        return self.plan.multiply(1.0, 0.0)

    def gemm_launch_stats(self):
        """(flops of this rank's GEMM launches in the last step, their summed device ms, number of launches)."""
        t = self.plan.gemm_times_ms()
        return self.plan.gemm_flops, sum(t), len(t)

    def mean_gemm_ms(self):
        t = self.plan.gemm_times_ms()
        return sum(t) / max(len(t), 1)

    def collectives(self, step_ms=None):
        """Device time and bus bandwidth of the collectives of the last step on this rank (bytes = (d-1)/d of the gathered /
        reduced buffer per rank), and -- given the step's duration -- how much of it the GEMM panels hid (overlapped plans run them on a
        second stream): exposed = step - (GEMM + accumulate time), hidden = communication time - exposed."""
        out = {}
        ov = self.plan.overlap()
        times = self.plan.op_times()
        for kind in ("allgather", "reduce"):
            sel = [(ms, wb) for k, ms, wb in times if k == kind]
            if not sel:
                continue
            ms, wb = sum(x[0] for x in sel), sum(x[1] for x in sel)
            out[kind] = {"count": len(sel), "ms": ms, "wire_bytes_per_rank": wb, "busbw_GBps": (wb / (ms * 1e-3) * 1e-9) if ms > 0 else None}
        out["overlapped"] = ov["enabled"]
        out["overlap_note"] = ov["why"]
        compute = sum(ms for k, ms, _ in times if k in ("gemm", "accumulate"))
        comm = sum(ms for k, ms, _ in times if k in ("allgather", "reduce"))
        out["gemm_launches"] = sum(1 for k, _, _ in times if k == "gemm")
        out["compute_ms"] = compute
        if step_ms is not None:
            exposed = max(0.0, step_ms - compute)
            out["exposed_ms"] = exposed
            out["hidden_ms"] = max(0.0, comm - exposed)
        return out

    def e2e(self, reps):
        """Same multiply through the host-pointer entry point: pinned local A, B in, local C out, per step."""
        import torch
        import torch.distributed as dist
        pl = self.plan
        multi = self.world > 1
        hA = torch.empty(max(pl.initial_elements[0], 1), dtype=self.tdt, pin_memory=True); hA[:pl.initial_elements[0]].copy_(pl.A.local)
        hB = torch.empty(max(pl.initial_elements[1], 1), dtype=self.tdt, pin_memory=True); hB[:pl.initial_elements[1]].copy_(pl.B.local)
        hC = torch.empty(max(pl.initial_elements[2], 1), dtype=self.tdt, pin_memory=True)
        pl.multiply_host(hA, hB, hC); torch.cuda.synchronize()
        if multi:
            dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            pl.multiply_host(hA, hB, hC)
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=self.device, dtype=torch.float64)
        if multi:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        n_cmp = min(pl.initial_elements[2], 1 << 20)
        ok = bool(torch.equal(hC[:n_cmp], pl.C.local[:n_cmp].cpu())) if n_cmp else True
        eb = hA.element_size()
        return {"value": self.flop_factor * pl.m * pl.n * pl.k / (ms.item() * 1e-3) * 1e-12, "unit": "TFLOP/s",
                "h2d_bytes_per_step": eb * (pl.initial_elements[0] + pl.initial_elements[1]), "d2h_bytes_per_step": eb * pl.initial_elements[2],
                "ms_per_step": ms.item(), "api": "cosma_b200_multiply_host (pinned host local A,B -> local C, per rank)",
                "matches_device_path": ok}

    def parity(self, samples=32):
        """One extra multiply on integer-valued operands defined by global coordinates (_formula): every rank checks, EXACTLY,
        (1) `samples` elements of its own part of C against dot products it evaluates itself in int64, and (2) the sum over its
        whole part of C against the identity sum_ij C_ij = sum_l (sum_i a_il)(sum_j b_lj) -- a checksum of checksums that covers
        every element the rank owns. The reference's own criterion (utils/cosma_utils.hpp:366-377) compares against a dense naive
        GEMM on rank 0, which does not exist at these sizes. -> dict for the bench line (ok = all ranks)."""
        import torch
        import torch.distributed as dist
        pl, dev, cplx = self.plan, self.device, self.cplx
        i64 = torch.int64

        def fill(mat, label, which):
            pos = 0
            for (r0, r1, c0, c1) in pl.local_blocks(label):
                nr = r1 - r0 + 1
                rows = torch.arange(r0, r1 + 1, device=dev, dtype=i64)[None, :]
                for cs in range(c0, c1 + 1, 512):                      # column chunks bound the int64 temporaries
                    ce = min(cs + 512, c1 + 1)
                    cols = torch.arange(cs, ce, device=dev, dtype=i64)[:, None]
                    re, im = _formula(which, rows, cols, cplx)         # (columns, rows): column-major order when flattened
                    cnt = nr * (ce - cs)
                    dst = mat.local[pos:pos + cnt]
                    if cplx:
                        dst.copy_(torch.complex(re.to(dst.real.dtype), im.to(dst.real.dtype)).reshape(-1))
                    else:
                        dst.copy_(re.to(dst.dtype).reshape(-1))
                    pos += cnt
        ok, checked = True, 0
        if not pl.idle:
            fill(pl.A, "A", 0)
            fill(pl.B, "B", 1)
            pl.C.local.fill_(float("nan"))
        pl.multiply(1.0, 0.0)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        k = pl.k
        if not pl.idle:
            gen = torch.Generator(); gen.manual_seed(4321 + self.rank)
            lk = torch.arange(k, device=dev, dtype=i64)
            pos = 0
            for (r0, r1, c0, c1) in pl.local_blocks("C"):
                nr, nc = r1 - r0 + 1, c1 - c0 + 1
                blk = pl.C.local[pos:pos + nr * nc]
                pos += nr * nc
                # (1) sampled elements
                for _ in range(samples):
                    li, lj = int(torch.randint(nr, (1,), generator=gen)), int(torch.randint(nc, (1,), generator=gen))
                    ar, ai = _formula(0, torch.tensor(r0 + li, device=dev, dtype=i64), lk, cplx)
                    br, bi = _formula(1, lk, torch.tensor(c0 + lj, device=dev, dtype=i64), cplx)
                    got = blk[lj * nr + li]
                    if cplx:
                        want = complex(int((ar * br - ai * bi).sum()), int((ar * bi + ai * br).sum()))
                        ok = ok and complex(got.item()) == want
                    else:
                        ok = ok and float(got.item()) == float(int((ar * br).sum()))
                    checked += 1
                # (2) checksum of the whole block
                rows = torch.arange(r0, r1 + 1, device=dev, dtype=i64)[:, None]
                cols = torch.arange(c0, c1 + 1, device=dev, dtype=i64)[None, :]
                tot_re = torch.zeros((), device=dev, dtype=i64); tot_im = torch.zeros((), device=dev, dtype=i64)
                step = max(1, min(k, (1 << 25) // max(nr, nc)))
                for l0 in range(0, k, step):
                    ls = torch.arange(l0, min(l0 + step, k), device=dev, dtype=i64)
                    ar, ai = _formula(0, rows, ls[None, :], cplx)       # (rows of the block, l)
                    br, bi = _formula(1, ls[:, None], cols, cplx)       # (l, columns of the block)
                    sar, sbr = ar.sum(0), br.sum(1)
                    if cplx:
                        sai, sbi = ai.sum(0), bi.sum(1)
                        tot_re += (sar * sbr - sai * sbi).sum(); tot_im += (sar * sbi + sai * sbr).sum()
                    else:
                        tot_re += (sar * sbr).sum()
                if cplx:
                    ok = ok and int(blk.real.to(torch.float64).to(i64).sum()) == int(tot_re) and int(blk.imag.to(torch.float64).to(i64).sum()) == int(tot_im)
                else:
                    ok = ok and bool(torch.isfinite(blk).all()) and int(blk.to(torch.float64).to(i64).sum()) == int(tot_re)
        flag = torch.tensor([1 if ok else 0, checked], device=dev, dtype=i64)
        if self.world > 1:
            ok_all = flag[:1].clone(); dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
            cnt = flag[1:].clone(); dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            ok, checked = bool(ok_all.item()), int(cnt.item())
        return {"ok": bool(ok), "exact": True, "sampled_elements": checked, "block_checksums": "every rank, every element of its C",
                "how": "integer-valued A, B from global coordinates; int64 dot products and sum_ij C_ij = sum_l (sum_i a_il)(sum_j b_lj)",
                "overlapped": pl.overlap()["enabled"]}

    def destroy(self):
        self.plan.destroy()
        if self.own_comm:
            self.comm.destroy()
