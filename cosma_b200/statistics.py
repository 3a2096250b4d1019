"""cosma_statistics for cosma_b200: what a multiply of (m, n, k) on P ranks will do, without running it (no GPU needed).

    python -m cosma_b200.statistics -m 32768 -n 32768 -k 32768 -P 8 [-s pm2,pn2,pk2] [-t double] [--layout]
    python -m cosma_b200.statistics -m 16384 -n 16384 -k 16384 -P 8 -t zdouble --pxgemm --block_a 256,256 -p 2,4 --transpose CN

Prints the strategy, the per-rank device arenas of the compiled schedule, the collectives with the bytes each rank puts on the
wire ((d-1)/d of the gathered / reduced buffer, ring size d) and the local GEMMs, plus a time estimate from this pool's measured
rates. The reference's tool of the same name (miniapp/cosma_statistics.cpp) walks its recursion with a counting communicator; here
the compiled op list already is that walk. --pxgemm adds what the relayouts of a p?gemm call on block-cyclic matrices move, for the
communication-optimal and the grid-adapted strategy, and what rank relabelling would keep in place (COSTA's comm_volume miniapp)."""
import argparse
import sys

from .distributed import MultiplyPlan

BYTES = {"float": ("s", 4), "double": ("d", 8), "zfloat": ("c", 8), "zdouble": ("z", 16)}
# measured on this pool (profiles/): DGEMM 36.8 TFLOP/s, 3xTF32 SGEMM ~155 TFLOP/s; NCCL over NVSwitch ~350 GB/s algorithm bandwidth
RATE_TFLOPS = {"d": 36.8, "z": 36.8, "s": 155.0, "c": 155.0}
WIRE_GBS = 350.0


def describe(m, n, k, P, steps="", dtype="double", ranks=None):
    code, eb = BYTES[dtype]
    ranks = list(range(P)) if ranks is None else ranks
    worst = {"arena": 0, "wire": 0, "flops": 0.0, "ops": 0}
    lines, strategy, P_used = [], None, None
    for r in ranks:
        pl = MultiplyPlan(None, m, n, k, steps, code, rank=r, nranks=P, allocate=False)
        strategy, P_used = pl.strategy, pl.P_used
        ops = pl.ops() if r < P_used else []
        wire = 0
        for op in ops:
            if op["kind"] != "gemm":
                total = sum(sum(p) for p in op["piece"])
                wire += (total - sum(op["piece"][op["my_pos"]])) * eb
        flops = sum((8.0 if code in "zc" else 2.0) * op["m"] * op["n"] * op["k"] for op in ops if op["kind"] == "gemm")
        worst["arena"] = max(worst["arena"], sum(pl.arena_elements) * eb)
        worst["wire"] = max(worst["wire"], wire)
        worst["flops"] = max(worst["flops"], flops)
        worst["ops"] = max(worst["ops"], len(ops))
        if r == ranks[0]:
            for op in ops:
                if op["kind"] == "gemm":
                    lines.append("    gemm       %d x %d x %d%s" % (op["m"], op["n"], op["k"], "" if op["beta"] == 0 else "  (accumulates)"))
                else:
                    total = sum(sum(p) for p in op["piece"])
                    lines.append("    %-10s %s  ring of %d, %d bucket(s), %.1f MB gathered/reduced, %s counts" %
                                 (op["kind"], "ABC"[op["matrix"]], len(op["ring"]), len(op["piece"][0]), total * eb / 1e6,
                                  "equal" if op["regular"] else "exact per-member"))
        pl.destroy()
    t_gemm = worst["flops"] / (RATE_TFLOPS[code] * 1e12)
    t_wire = worst["wire"] / (WIRE_GBS * 1e9)
    return {"strategy": strategy, "P_used": P_used, "lines": lines, "arena_bytes": worst["arena"], "wire_bytes": worst["wire"],
            "flops": worst["flops"], "ops": worst["ops"], "t_gemm_ms": t_gemm * 1e3, "t_wire_ms": t_wire * 1e3}


def layouts(m, n, k, P, steps=""):
    """COSMA's native layout of A, B, C as grids (the reference's miniapp/layout_miniapp.cpp prints the same): row split points, column split
    points and the owner rank of every block."""
    from . import planning
    full = steps or planning.strategy(m, n, k, P)[0]
    P_used = planning.strategy(m, n, k, P, 0, full)[1] if full else 1
    out = {}
    for label, (rows, cols) in (("A", (m, k)), ("B", (k, n)), ("C", (m, n))):
        per = planning.mapper_layout(label, m, n, k, P_used, full)
        rs = sorted({b[0] for bl in per for b in bl} | {rows})
        cs = sorted({b[2] for bl in per for b in bl} | {cols})
        owners = [[-1] * (len(cs) - 1) for _ in range(len(rs) - 1)]
        for r, bl in enumerate(per):
            for (r0, r1, c0, c1) in bl:
                owners[rs.index(r0)][cs.index(c0)] = r
        out[label] = (rs, cs, owners)
    return full, out


def pxgemm_relayout(m, n, k, P, nprow, npcol, order, trans, blocks, steps=""):
    """What the relayouts of a p?gemm call move: for op(A), op(B) (block-cyclic -> COSMA's native layout of `steps`) and C (native ->
    block-cyclic) the elements that stay on their rank and those that cross, plus what rank relabelling (costa::optimal_reordering on
    the summed volume graph, reference cosma_pxgemm.cpp:255-271) would keep in place. blocks: ((mb, nb) of A, of B, of C)."""
    from . import costa, planning
    ta, tb = trans[0].upper(), trans[1].upper()
    shapes = {"A": (m, k) if ta == "N" else (k, m), "B": (k, n) if tb == "N" else (n, k), "C": (m, n)}
    full, native = layouts(m, n, k, P, steps)
    total = [[0] * P for _ in range(P)]
    rows = []
    for label, blk, tr in (("A", blocks[0], ta), ("B", blocks[1], tb), ("C", blocks[2], "N")):
        r, c = shapes[label]
        rs, cs, ow, _ = costa.scalapack_grid(max(r, 1), r, c, 1, 1, r, c, blk[0], blk[1], nprow, npcol, order, 0, 0, "C", 0)
        user = (list(rs), list(cs), [[int(ow[i][j]) for j in range(len(cs) - 1)] for i in range(len(rs) - 1)])
        nat = native[label]
        nat = (nat[0], nat[1], [[max(o, 0) for o in row] for row in nat[2]])
        vol = planning.comm_volume(user, nat, tr, P) if label != "C" else planning.comm_volume(nat, user, "N", P)
        stay = sum(vol[u][u] for u in range(P))
        rows.append((label, r * c, stay, r * c - stay))
        for u in range(P):
            for v in range(P):
                total[u][v] += vol[u][v]
    perm, reordered = planning.optimal_reordering(total)
    kept = sum(total[u][u] for u in range(P))
    kept_relabelled = sum(total[min(u, perm[u])][max(u, perm[u])] for u in range(P) if perm[u] >= u)
    return {"strategy": full, "matrices": rows, "stay": kept, "stay_relabelled": kept_relabelled if reordered else kept, "permutation": perm,
            "reordered": reordered, "elements": sum(x[1] for x in rows)}


def main(argv=None):
    ap = argparse.ArgumentParser(prog="cosma_b200.statistics")
    ap.add_argument("-m", type=int, required=True)
    ap.add_argument("-n", type=int, required=True)
    ap.add_argument("-k", type=int, required=True)
    ap.add_argument("-P", type=int, required=True)
    ap.add_argument("-s", "--steps", default="")
    ap.add_argument("-t", "--type", default="double", choices=sorted(BYTES))
    ap.add_argument("--layout", action="store_true", help="also print the native layout grids of A, B, C (layout_miniapp)")
    ap.add_argument("--pxgemm", action="store_true", help="also print what the relayouts of a p?gemm call on block-cyclic matrices move")
    ap.add_argument("--block_a", default="128,128")
    ap.add_argument("--block_b", default="128,128")
    ap.add_argument("--block_c", default="128,128")
    ap.add_argument("-p", "--p_grid", default="", help="nprow,npcol (default: the most square grid of P)")
    ap.add_argument("--order", default="R", choices=["R", "C"])
    ap.add_argument("--transpose", default="NN")
    a = ap.parse_args(argv)
    ranks = None if a.P <= 64 else [0, 1, a.P // 2, a.P - 1]
    d = describe(a.m, a.n, a.k, a.P, a.steps, a.type, ranks)
    print("problem   : %d x %d x %d, %s, %d rank(s) (%d used)" % (a.m, a.n, a.k, a.type, a.P, d["P_used"]))
    print("strategy  : %s" % (d["strategy"] or "(single GEMM)"))
    print("schedule of rank 0 (%d ops):" % len(d["lines"]))
    for ln in d["lines"][:40]:
        print(ln)
    if len(d["lines"]) > 40:
        print("    ... %d more" % (len(d["lines"]) - 40))
    print("per rank (worst): device arenas %.2f GB, wire %.1f MB, GEMM %.2f TFLOP" % (d["arena_bytes"] / 1e9, d["wire_bytes"] / 1e6, d["flops"] / 1e12))
    print("estimate  : GEMM %.2f ms + collectives %.2f ms (not overlapped) -> %.1f %% of the step is communication" %
          (d["t_gemm_ms"], d["t_wire_ms"], 100.0 * d["t_wire_ms"] / max(d["t_gemm_ms"] + d["t_wire_ms"], 1e-12)))
    if a.P > 1 and d["strategy"]:
        # can the host-memory entry point run this schedule as pipelined column panels (COSMA_B200_HOST_PANELS)?
        import ctypes
        code = BYTES[a.type][0]
        ok_for = []
        for c in (2, 4, 8):
            verdicts = []
            for r in range(d["P_used"]):
                pl = MultiplyPlan(None, a.m, a.n, a.k, a.steps, code, rank=r, nranks=a.P, allocate=False)
                nb, nc, ok = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
                pl.lib.cosma_b200_plan_host_panel(pl.handle, c, 0, None, 0, ctypes.byref(nb), None, 0, ctypes.byref(nc), ctypes.byref(ok))
                n_gemm = sum(op["kind"] == "gemm" for op in pl.ops())
                verdicts.append(bool(ok.value) and n_gemm == 1)
                pl.destroy()
                if not verdicts[-1] or r >= 7:
                    break
            if verdicts and all(verdicts):
                ok_for.append(c)
        print("host panels: %s" % ("local B / C can be cut into %s column panels (cosma_b200_plan_host_panel)" % " / ".join(map(str, ok_for)) if ok_for
                                   else "layout cannot be cut into column panels"))
    if a.layout:
        _, grids = layouts(a.m, a.n, a.k, a.P, a.steps)
        for label in "ABC":
            rs, cs, owners = grids[label]
            print("layout of %s: row splits %s\n             col splits %s" % (label, rs, cs))
            for row in owners[:16]:
                print("             owners " + " ".join("%3d" % o for o in row[:32]) + (" ..." if len(row) > 32 else ""))
            if len(owners) > 16:
                print("             ... %d more block rows" % (len(owners) - 16))
    if a.pxgemm:
        from . import planning
        pair = lambda t: tuple(int(x) for x in t.split(","))
        if a.p_grid:
            nprow, npcol = pair(a.p_grid)
        else:
            nprow = max(d for d in range(1, int(a.P ** 0.5) + 1) if a.P % d == 0)
            npcol = a.P // nprow
        blocks = (pair(a.block_a), pair(a.block_b), pair(a.block_c))
        ta, tb = a.transpose[0].upper(), a.transpose[1].upper()
        shape = lambda r, c, t: (r, c) if t == "N" else (c, r)
        desc = lambda rc, blk: [1, 0, rc[0], rc[1], blk[0], blk[1], 0, 0, max(rc[0], 1)]
        variants = [("communication-optimal strategy", a.steps)]
        prefix = planning.adapt_strategy(a.m, a.n, a.k, a.P, desc(shape(a.m, a.k, ta), blocks[0]), 1, 1, desc(shape(a.k, a.n, tb), blocks[1]), 1, 1,
                                         desc((a.m, a.n), blocks[2]), 1, 1, ta, tb, nprow, npcol, a.order)
        if prefix and not a.steps:
            variants.append(("COSMA_ADAPT_STRATEGY=ON", planning.strategy(a.m, a.n, a.k, a.P, 0, prefix)[0]))
        print("p?gemm    : grid %d x %d (%s), op = %s, blocks A %s B %s C %s" % (nprow, npcol, a.order, a.transpose.upper(), blocks[0], blocks[1], blocks[2]))
        for name, steps in variants:
            r = pxgemm_relayout(a.m, a.n, a.k, a.P, nprow, npcol, a.order, a.transpose, blocks, steps)
            print("  %s [%s]" % (name, r["strategy"] if len(r["strategy"]) < 60 else r["strategy"][:57] + "..."))
            for label, elems, stay, cross in r["matrices"]:
                print("    relayout of %s: %5.1f %% of %d elements change rank" % (label, 100.0 * cross / max(elems, 1), elems))
            print("    in place: %.1f %%; with rank relabelling (COSMA_B200_REORDER_RANKS=ON): %.1f %%%s" %
                  (100.0 * r["stay"] / max(r["elements"], 1), 100.0 * r["stay_relabelled"] / max(r["elements"], 1),
                   "" if r["reordered"] else " (nothing to gain)"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
