#!/usr/bin/env python3
"""Generates tests/golden/ref_multirank_*.npz from the UNMODIFIED reference (oracle/_ref/ref_driver on minimpi ranks):
inputs, strategy and every rank's raw local C buffer of cosma::multiply. Run in the build container (needs
/root/reference to have been compiled by `make -C oracle ref`); the fixtures travel, the reference does not."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402

CASES = [  # (m, n, k, P, steps, dtype, alpha, beta)
    (8, 4, 2, 4, "pm2,sm2,pn2", "d", 1.0, 1.0),          # tests/mapper.cpp:408-562 layout golden, here with values
    (30, 35, 40, 4, "", "d", 1.0, 1.0),
    (20, 20, 20, 3, "sk2,pm3", "d", 1.0, 1.0),
    (16, 16, 16, 16, "pm2,pn2,pk2,pm2", "d", 2.0, 0.0),
    (100, 100, 100, 12, "pm2,pn2,pk3", "z", 1.0 - 2.0j, 0.5j),
    (100, 100, 100, 8, "sm2,pn2,sk2,pm2,sn2,pk2", "s", 1.0, -1.0),  # tests/scalar_matmul.cpp
    (64, 64, 64, 8, "pm2,pn2,pk2", "d", 1.0, 0.0),       # the BASELINE 32768^3 P=8 strategy in miniature
    (32, 32, 512, 8, "pk8", "d", 1.0, 0.0),              # the BASELINE large-K strategy in miniature
    (96, 96, 96, 2, "pk2", "d", 1.0, 0.0),               # BASELINE configs[0] / 32768^3 at P=2 in miniature
    (60, 52, 44, 2, "sm2,sk3,pn2", "z", 2.0, 1.0),     # (a strategy ENDING in a sequential step makes the reference divide by zero)
    (64, 64, 64, 4, "pn2,pk2", "d", 1.0, 0.0),           # the BASELINE 32768^3 P=4 strategy in miniature
]


def main():
    out = os.path.join(ROOT, "tests", "golden")
    for i, (m, n, k, P, steps, dtype, alpha, beta) in enumerate(CASES):
        rng = np.random.default_rng(1000 + i)
        def one(r, c):
            v = rng.integers(-4, 6, size=(r, c)).astype(np.float64)
            return v + 1j * rng.integers(-4, 6, size=(r, c)) if dtype in "cz" else v
        A, B, C = one(m, k), one(k, n), one(m, n)
        locs, _ = o.ref_multiply_ranks(dtype, m, n, k, P, steps, alpha, beta, A, B, C)
        small = np.int16 if dtype in "sd" else None
        data = dict(m=m, n=n, k=k, P=P, steps=steps, dtype=dtype, alpha=complex(alpha), beta=complex(beta),
                    A=A.astype(o.NPDT[dtype]), B=B.astype(o.NPDT[dtype]), C=C.astype(o.NPDT[dtype]))
        for r, loc in enumerate(locs):
            if loc is not None:
                data["local_c_%d" % r] = loc
        name = "ref_multirank_%02d_%dx%dx%d_P%d.npz" % (i, m, n, k, P)
        np.savez_compressed(os.path.join(out, name), **data)
        print(name, os.path.getsize(os.path.join(out, name)), "bytes")


if __name__ == "__main__":
    main()
