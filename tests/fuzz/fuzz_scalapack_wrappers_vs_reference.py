"""Randomised parity of the ScaLAPACK-style entry points against the UNMODIFIED reference wrappers on several ranks:
  p?tran / p?tranu / p?tranc and p?gemr2d: OUR transform plans (cosma_b200_scalapack_layout + cosma_b200_transform_plan_create) run in CPU
      lock-step; every rank's local array must equal the reference's (costa::pxtran_op / costa::pxgemr2d) BIT FOR BIT, padding included;
  p?gemm: our three-phase pipeline (tests/test_pxgemm_cpu.py) against cosma::pxgemm of the reference, exact on integer matrices.
Random matrix sizes, block sizes, process grids (both numberings), sub-matrix origins, rsrc/csrc, transposes, alpha/beta.
    python tests/fuzz/fuzz_scalapack_wrappers_vs_reference.py SEED N"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import costa_sim as sim  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from test_pxgemm_cpu import run_pdgemm_on_cpu  # noqa: E402
from test_pxtran_cpu import _simulate  # noqa: E402


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rnd = random.Random(seed)
    orc.ref(); orc.lib()
    bad = ran = 0
    for it in range(N):
        nprow, npcol = rnd.choice([(1, 1), (1, 2), (2, 1), (2, 2), (2, 3), (3, 2), (1, 4), (4, 2), (2, 4)])
        order = rnd.choice("RC")
        P = nprow * npcol
        kind = rnd.choice(["tran", "gemr2d", "gemm"])
        rng = np.random.default_rng(seed * 100003 + it)
        blk = lambda: (rnd.randint(1, 12), rnd.randint(1, 12))
        if kind in ("tran", "gemr2d"):
            dtype = rnd.choice("dzsc") if kind == "tran" else rnd.choice("dzs")
            op = "N" if kind == "gemr2d" else (rnd.choice("TC") if dtype in "zc" else "T")
            m, n = rnd.randint(1, 70), rnd.randint(1, 70)
            (ia, ja), (ic, jc) = (rnd.randint(1, 6), rnd.randint(1, 6)), (rnd.randint(1, 6), rnd.randint(1, 6))
            extra = rnd.randint(0, 9)
            alpha, beta = (1.0, 0.0) if kind == "gemr2d" else rnd.choice([(1.0, 0.0), (2.0, -1.0), (1.0, 1.0), (0.5, 0.0)])
            if dtype in "cz" and kind == "tran":
                alpha = alpha * (1 - 0.5j)
            am, an = (m, n) if op == "N" else (n, m)
            GA = sim.random_values(rng, (am + ia - 1 + extra, an + ja - 1 + extra), dtype)
            GC = sim.random_values(rng, (m + ic - 1 + extra, n + jc - 1 + extra), dtype)
            orderc = rnd.choice("RC") if kind == "gemr2d" else order
            src = (rnd.randrange(nprow), rnd.randrange(npcol)) if kind == "tran" else (0, 0)
            bcA = sim.BlockCyclic(GA.shape[0], GA.shape[1], *blk(), nprow, npcol, order, src[0], src[1], lld_pad=rnd.randint(0, 2))
            bcC = sim.BlockCyclic(GC.shape[0], GC.shape[1], *blk(), nprow, npcol, orderc, 0, 0, lld_pad=rnd.randint(0, 2))
            a_loc = [bcA.scatter(GA, r, fill=77) for r in range(P)]
            c_loc = [bcC.scatter(GC, r, fill=55) for r in range(P)]
            try:
                if kind == "tran":
                    want = orc.ref_pxtran_ranks(dtype, order, nprow, npcol, op, m, n, alpha, a_loc, ia, ja, [bcA.desc(r) for r in range(P)], beta,
                                                [x.copy() for x in c_loc], ic, jc, [bcC.desc(r) for r in range(P)])
                else:
                    want = orc.ref_pxgemr2d_ranks(dtype, order, nprow, npcol, m, n, a_loc, ia, ja, [bcA.desc(r) for r in range(P)],
                                                  [x.copy() for x in c_loc], ic, jc, [bcC.desc(r) for r in range(P)], orderc=orderc)
            except Exception as e:
                print("reference failed:", kind, dtype, op, m, n, (nprow, npcol, order), str(e)[:80])
                continue
            _simulate(orc, dtype, op, m, n, alpha, beta, bcA, bcC, a_loc, c_loc, ia, ja, ic, jc, P)
            ok = all(np.array_equal(c_loc[r].view(np.uint8), want[r].view(np.uint8)) for r in range(P))
            desc = (kind, dtype, op, m, n, (nprow, npcol, order, orderc), (ia, ja, ic, jc), extra, alpha, beta)
        else:
            ta, tb = rnd.choice("NT"), rnd.choice("NT")
            m, n, k = rnd.randint(1, 60), rnd.randint(1, 60), rnd.randint(1, 60)
            sub = [(rnd.randint(1, 5), rnd.randint(1, 5)) for _ in range(3)]
            extra = rnd.randint(0, 7)
            am, an = (m, k) if ta == "N" else (k, m)
            bm, bn = (k, n) if tb == "N" else (n, k)
            dims = [(am, an), (bm, bn), (m, n)]
            blks = [blk() for _ in range(3)]
            alpha, beta = rnd.choice([(1.0, 0.0), (2.0, -1.0), (1.0, 1.0), (0.5, 0.5)])
            c = dict(m=m, n=n, k=k, ta=ta, tb=tb, alpha=alpha, beta=beta, p_rows=nprow, p_cols=npcol, order=order)
            for x, nm in enumerate("abc"):
                c["m" + nm], c["n" + nm] = dims[x][0] + sub[x][0] - 1 + extra, dims[x][1] + sub[x][1] - 1 + extra
                c["bm" + nm], c["bn" + nm] = blks[x]
                c["i" + nm], c["j" + nm] = sub[x]
                c["src_m" + nm], c["src_n" + nm] = rnd.randrange(nprow), rnd.randrange(npcol)
            got, want_dense, _ = run_pdgemm_on_cpu(orc, c, seed * 7919 + it)
            # the reference on the same data (run_pdgemm_on_cpu draws G with default_rng(seed) in the same order)
            rng2 = np.random.default_rng(seed * 7919 + it)
            shapes = [(c["ma"], c["na"]), (c["mb"], c["nb"]), (c["mc"], c["nc"])]
            G = [sim.random_values(rng2, s, "d") for s in shapes]
            bc = [sim.BlockCyclic(shapes[x][0], shapes[x][1], blks[x][0], blks[x][1], nprow, npcol, order, c["src_m" + "abc"[x]], c["src_n" + "abc"[x]])
                  for x in range(3)]
            locs = [[bc[x].scatter(G[x], r) for r in range(P)] for x in range(3)]
            descs = [[bc[x].desc(r) for r in range(P)] for x in range(3)]
            try:
                outs, _ = orc.ref_pxgemm_ranks("d", order, nprow, npcol, ta, tb, m, n, k, alpha, locs[0], sub[0][0], sub[0][1], descs[0], locs[1], sub[1][0],
                                               sub[1][1], descs[1], beta, locs[2], sub[2][0], sub[2][1], descs[2])
            except Exception as e:
                print("reference failed: gemm", m, n, k, (nprow, npcol, order), str(e)[:80])
                continue
            refg = np.zeros_like(G[2])
            for r in range(P):
                bc[2].gather_into(refg, outs[r], r)
            ok = np.allclose(got, want_dense, rtol=1e-14, atol=0) and np.allclose(refg, got, rtol=1e-14, atol=0)
            desc = ("gemm", ta, tb, m, n, k, (nprow, npcol, order), sub, extra, alpha, beta)
        ran += 1
        if not ok:
            bad += 1
            print("MISMATCH", desc)
    print("cases run %d, mismatches %d" % (ran, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
