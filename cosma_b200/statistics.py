"""cosma_statistics for cosma_b200: what a multiply of (m, n, k) on P ranks will do, without running it (no GPU needed).

    python -m cosma_b200.statistics -m 32768 -n 32768 -k 32768 -P 8 [-s pm2,pn2,pk2] [-t double]

Prints the strategy, the per-rank device arenas of the compiled schedule, the collectives with the bytes each rank puts on the
wire ((d-1)/d of the gathered / reduced buffer, ring size d) and the local GEMMs, plus a time estimate from this pool's measured
rates. The reference's tool of the same name (miniapp/cosma_statistics.cpp) walks its recursion with a counting communicator; here
the compiled op list already is that walk."""
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


def main(argv=None):
    ap = argparse.ArgumentParser(prog="cosma_b200.statistics")
    ap.add_argument("-m", type=int, required=True)
    ap.add_argument("-n", type=int, required=True)
    ap.add_argument("-k", type=int, required=True)
    ap.add_argument("-P", type=int, required=True)
    ap.add_argument("-s", "--steps", default="")
    ap.add_argument("-t", "--type", default="double", choices=sorted(BYTES))
    ap.add_argument("--layout", action="store_true", help="also print the native layout grids of A, B, C (layout_miniapp)")
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
    if a.layout:
        _, grids = layouts(a.m, a.n, a.k, a.P, a.steps)
        for label in "ABC":
            rs, cs, owners = grids[label]
            print("layout of %s: row splits %s\n             col splits %s" % (label, rs, cs))
            for row in owners[:16]:
                print("             owners " + " ".join("%3d" % o for o in row[:32]) + (" ..." if len(row) > 32 else ""))
            if len(owners) > 16:
                print("             ... %d more block rows" % (len(owners) - 16))
    return 0


if __name__ == "__main__":
    sys.exit(main())
