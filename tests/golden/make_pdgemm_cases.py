"""Extracts the 50 parameter sets of the reference's ScaLAPACK-wrapper test (tests/pdgemm.cpp: INSTANTIATE_TEST_CASE_P with
cosma::pxgemm_params<double>{...}) into tests/golden/pdgemm_cases.json, expanding the short 12-argument form with the rules of
pxgemm_params::initialize (src/cosma/pxgemm_params.hpp:112-205: blocks of op(A)/op(B) from (bm, bn, bk), sub-matrices at (1, 1),
row-major grid, lld = max_leading_dimension, sources (0, 0)). Run in the build container (needs /root/reference); the GPU box only
sees the JSON."""
import json
import os
import re

REF = "/root/reference/tests/pdgemm.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pdgemm_cases.json")


def max_lld(n, nb, p):
    """cosma::scalapack::max_leading_dimension (src/cosma/scalapack.cpp:105-119)"""
    whole = n // nb
    return whole // p * nb + ((n % nb) if whole % p == 0 else nb)


def main():
    body = re.sub(r"//[^\n]*", "", open(REF).read())
    cases = []
    for m in re.finditer(r"pxgemm_params<double>\s*\{([^}]*)\}", body):
        t = [x.strip().strip("'") for x in m.group(1).replace("\n", " ").split(",") if x.strip()]
        if len(t) == 37:
            (ma, na, mb, nb, mc, nc, bma, bna, bmb, bnb, bmc, bnc, ia, ja, ib, jb, ic, jc, mm, nn, kk) = [int(x) for x in t[:21]]
            ta, tb = t[21], t[22]
            alpha, beta = float(t[23]), float(t[24])
            lld = [int(x) for x in t[25:28]]
            prow, pcol, order = int(t[28]), int(t[29]), t[30]
            src = [int(x) for x in t[31:37]]
        elif len(t) == 12:
            mm, nn, kk, bm, bn, bk, prow, pcol = [int(x) for x in t[:8]]
            ta, tb = t[8], t[9]
            alpha, beta = float(t[10]), float(t[11])
            tr = lambda flag, row, col: row if flag != "N" else col
            ma, na = tr(ta, kk, mm), tr(ta, mm, kk)
            mb, nb = tr(tb, nn, kk), tr(tb, kk, nn)
            mc, nc = mm, nn
            bma, bna = tr(ta, bk, bm), tr(ta, bm, bk)
            bmb, bnb = tr(tb, bn, bk), tr(tb, bk, bn)
            bmc, bnc = bm, bn
            ia = ja = ib = jb = ic = jc = 1
            order = "R"
            lld = [max_lld(ma, bma, prow), max_lld(mb, bmb, prow), max_lld(mc, bmc, prow)]
            src = [0] * 6
        else:
            raise SystemExit("unexpected parameter count %d" % len(t))
        cases.append(dict(ma=ma, na=na, mb=mb, nb=nb, mc=mc, nc=nc, bma=bma, bna=bna, bmb=bmb, bnb=bnb, bmc=bmc, bnc=bnc, ia=ia, ja=ja, ib=ib, jb=jb,
                          ic=ic, jc=jc, m=mm, n=nn, k=kk, ta=ta.upper(), tb=tb.upper(), alpha=alpha, beta=beta, lld_a=lld[0], lld_b=lld[1], lld_c=lld[2],
                          p_rows=prow, p_cols=pcol, order=order.upper(), src_ma=src[0], src_na=src[1], src_mb=src[2], src_nb=src[3], src_mc=src[4],
                          src_nc=src[5]))
    with open(OUT, "w") as f:
        json.dump({"source": "eth-cscs/COSMA tests/pdgemm.cpp (INSTANTIATE_TEST_CASE_P Default)", "cases": cases}, f, indent=0)
    # the same cases as one line of 37 tokens each, for the C++ program tests/cpp/test_pxgemm.cpp
    keys = ["ma", "na", "mb", "nb", "mc", "nc", "bma", "bna", "bmb", "bnb", "bmc", "bnc", "ia", "ja", "ib", "jb", "ic", "jc", "m", "n", "k", "ta", "tb",
            "alpha", "beta", "lld_a", "lld_b", "lld_c", "p_rows", "p_cols", "order", "src_ma", "src_na", "src_mb", "src_nb", "src_mc", "src_nc"]
    with open(OUT.replace(".json", ".txt"), "w") as f:
        for c in cases:
            f.write(" ".join(str(c[k]) for k in keys) + "\n")
    print(len(cases), "cases ->", OUT)


if __name__ == "__main__":
    main()
