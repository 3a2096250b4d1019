"""Host planning layer through the C ABI: Strategy and Mapper (mirrors cosma::Strategy / cosma::Mapper)."""
import ctypes

from . import _lib


def strategy(m, n, k, P, mem_limit=0, prefix=""):
    """-> (steps string e.g. 'pm2,pn2,pk2', ranks actually used, elements of memory per rank)"""
    lib = _lib.load()
    out = ctypes.create_string_buffer(4096)
    P_out = ctypes.c_int(0)
    mem = ctypes.c_longlong(0)
    st = lib.cosma_b200_strategy(ctypes.c_int(m), ctypes.c_int(n), ctypes.c_int(k), ctypes.c_int(P),
                                 ctypes.c_longlong(mem_limit), prefix.encode(), out, ctypes.c_int(4096),
                                 ctypes.byref(P_out), ctypes.byref(mem))
    _lib.check(st, "cosma_b200_strategy")
    return out.value.decode(), P_out.value, mem.value


def mapper_layout(label, m, n, k, P, steps):
    """-> per rank, list of (row_first, row_last, col_first, col_last) blocks in local-buffer order"""
    lib = _lib.load()
    counts = (ctypes.c_int * max(P, 1))()
    cap = 4 * 65536
    out = (ctypes.c_int * cap)()
    total = ctypes.c_int(0)
    st = lib.cosma_b200_mapper_layout(ctypes.c_char(label.encode()), ctypes.c_int(m), ctypes.c_int(n), ctypes.c_int(k),
                                      ctypes.c_int(P), steps.encode(), counts, out, ctypes.c_int(cap), ctypes.byref(total))
    _lib.check(st, "cosma_b200_mapper_layout")
    res, pos = [], 0
    for r in range(P):
        blocks = []
        for _ in range(counts[r]):
            blocks.append(tuple(out[4 * pos:4 * pos + 4]))
            pos += 1
        res.append(blocks)
    return res


def local_coordinates(label, m, n, k, P, steps, gi, gj):
    lib = _lib.load()
    li = ctypes.c_int64(0)
    rk = ctypes.c_int(0)
    st = lib.cosma_b200_mapper_local_coordinates(ctypes.c_char(label.encode()), ctypes.c_int(m), ctypes.c_int(n),
                                                 ctypes.c_int(k), ctypes.c_int(P), steps.encode(), ctypes.c_int(gi),
                                                 ctypes.c_int(gj), ctypes.byref(li), ctypes.byref(rk))
    _lib.check(st, "cosma_b200_mapper_local_coordinates")
    return li.value, rk.value


def global_coordinates(label, m, n, k, P, steps, local_idx, rank):
    lib = _lib.load()
    gi = ctypes.c_int(0)
    gj = ctypes.c_int(0)
    st = lib.cosma_b200_mapper_global_coordinates(ctypes.c_char(label.encode()), ctypes.c_int(m), ctypes.c_int(n),
                                                  ctypes.c_int(k), ctypes.c_int(P), steps.encode(), ctypes.c_int64(local_idx),
                                                  ctypes.c_int(rank), ctypes.byref(gi), ctypes.byref(gj))
    _lib.check(st, "cosma_b200_mapper_global_coordinates")
    return gi.value, gj.value


def fit_strategy(m, n, k, P, elem_bytes, budget_bytes, prefix=""):
    """cosma_b200_fit_strategy: the automatic strategy tightened with sequential steps until the arenas of its compiled schedule fit
    budget_bytes of device memory per rank. -> (steps, ranks used, footprint bytes). Raises when nothing fits."""
    lib = _lib.load()
    out = ctypes.create_string_buffer(8192)
    P_out, foot = ctypes.c_int(0), ctypes.c_longlong(0)
    st = lib.cosma_b200_fit_strategy(ctypes.c_int(m), ctypes.c_int(n), ctypes.c_int(k), ctypes.c_int(P), prefix.encode(), ctypes.c_int(elem_bytes),
                                     ctypes.c_longlong(budget_bytes), out, ctypes.c_int(8192), ctypes.byref(P_out), ctypes.byref(foot))
    _lib.check(st, "cosma_b200_fit_strategy")
    return out.value.decode(), P_out.value, foot.value


def adapt_strategy(m, n, k, P, desca, ia, ja, descb, ib, jb, descc, ic, jc, transa, transb, nprow, npcol, order):
    """cosma_b200_adapt_strategy: the strategy prefix that mirrors the block-cyclic grid of the largest operand ("" if the reference's
    conditions do not hold). desc*: 9-int ScaLAPACK descriptors (sequences)."""
    lib = _lib.load()
    arr = [(ctypes.c_int * 9)(*[int(x) for x in d]) for d in (desca, descb, descc)]
    out = ctypes.create_string_buffer(512)
    st = lib.cosma_b200_adapt_strategy(m, n, k, P, arr[0], ia, ja, arr[1], ib, jb, arr[2], ic, jc, ctypes.c_char(transa.encode()),
                                       ctypes.c_char(transb.encode()), nprow, npcol, ctypes.c_char(order.encode()), out, 512)
    _lib.check(st, "cosma_b200_adapt_strategy")
    return out.value.decode()


def comm_volume(grid_a, grid_b, trans, n_ranks):
    """cosma_b200_comm_volume: elements exchanged between every pair of ranks when a matrix moves from grid_a (transposed first when trans
    != 'N') to grid_b. grid = (rowsplit, colsplit, owners[rows][cols]). -> n_ranks x n_ranks nested list, upper triangle, [u][u] stays."""
    lib = _lib.load()

    def pack(g):
        rs, cs, ow = g
        flat = [int(o) for row in ow for o in row]
        return len(rs) - 1, len(cs) - 1, (ctypes.c_int * len(rs))(*rs), (ctypes.c_int * len(cs))(*cs), (ctypes.c_int * max(len(flat), 1))(*flat)
    a, b = pack(grid_a), pack(grid_b)
    vol = (ctypes.c_longlong * (n_ranks * n_ranks))()
    st = lib.cosma_b200_comm_volume(a[0], a[1], a[2], a[3], a[4], b[0], b[1], b[2], b[3], b[4], ctypes.c_char(trans.encode()), n_ranks, vol)
    _lib.check(st, "cosma_b200_comm_volume")
    return [[vol[u * n_ranks + v] for v in range(n_ranks)] for u in range(n_ranks)]


def optimal_reordering(volume):
    """cosma_b200_optimal_reordering on a comm_volume() result (or a sum of several). -> (permutation, reordered)."""
    lib = _lib.load()
    n = len(volume)
    flat = (ctypes.c_longlong * (n * n))(*[int(volume[u][v]) for u in range(n) for v in range(n)])
    perm, flag = (ctypes.c_int * n)(), ctypes.c_int(0)
    _lib.check(lib.cosma_b200_optimal_reordering(n, flat, perm, ctypes.byref(flag)), "cosma_b200_optimal_reordering")
    return list(perm), bool(flag.value)
