"""Shared problem lists for the parity tests."""

# The 40 (m, n, k, P, strategy) cases of the reference's distributed test (tests/multiply.cpp:142-321), with the
# (divisors, split dimensions, step types) triplets written in the miniapp's "-s" notation; "" = automatic strategy.
def _s(divs, dims, types):
    return ",".join("%s%s%d" % (t, d, v) for v, d, t in zip(divs, dims, types))


REFERENCE_MULTIPLY_CASES = [
    (4, 4, 4, 1, ""), (3, 4, 5, 1, ""),
    (8, 4, 2, 4, _s([2, 2, 2], "mmn", "psp")), (8, 4, 2, 4, ""),
    (4, 4, 4, 2, _s([2], "m", "p")), (4, 4, 4, 2, ""),
    (4, 4, 4, 4, _s([2, 2, 2], "mnn", "spp")),
    (30, 35, 40, 4, ""),
    (8, 8, 2, 2, _s([2, 2, 2], "mmn", "ssp")), (8, 8, 2, 2, ""),
    (16, 4, 4, 4, _s([2, 2], "mm", "pp")), (16, 4, 4, 4, ""),
    (20, 20, 20, 3, _s([2, 3], "km", "sp")), (20, 20, 20, 3, ""),
    (16, 16, 16, 16, _s([2, 2, 2, 2], "mnkm", "pppp")), (16, 16, 16, 16, ""),
    (20, 30, 25, 4, _s([2, 2, 2, 2], "mnkm", "sspp")), (20, 30, 25, 4, ""),
    (100, 100, 100, 10, _s([2, 2, 2, 5], "mnkm", "spsp")), (100, 100, 100, 10, ""),
    (4, 4, 5, 4, _s([2, 2, 2, 2], "mnkm", "spsp")), (4, 4, 5, 4, ""),
    (10, 10, 10, 12, _s([2, 2, 3], "mnk", "ppp")),
    (100, 100, 100, 12, _s([2, 2, 3], "mnk", "ppp")), (100, 100, 100, 12, ""),
    (100, 100, 100, 4, ""),
    (100, 100, 100, 7, _s([7], "m", "p")), (100, 100, 100, 7, ""),
    (100, 100, 100, 8, _s([2] * 6, "mnkmnk", "spspsp")), (100, 100, 100, 8, ""),
    (100, 100, 100, 4, _s([2, 2], "mk", "pp")),
    (100, 100, 100, 8, _s([2, 2], "mk", "pp")),
    (100, 100, 100, 8, _s([2] * 6, "mknnmk", "sssppp")),
    (100, 100, 100, 8, _s([2] * 6, "kmnkmn", "spspsp")),
    (200, 200, 200, 8, _s([3, 3, 3, 2, 2, 2], "kmnknm", "sssppp")), (200, 200, 200, 8, ""),
    (200, 200, 200, 8, _s([3, 2, 3, 2, 3, 2], "mnkmnk", "spspsp")),
    (512, 32, 736, 8, _s([2, 2, 2], "kmk", "ppp")),
]

# tests/scalar_matmul.cpp:7-39
SCALAR_MATMUL_CASE = (100, 100, 100, 8, "sm2,pn2,sk2,pm2,sn2,pk2")

# BASELINE.json configs (automatic strategy) at every GPU count the bench uses, plus large/irregular extras
BASELINE_CASES = [
    (2000, 2000, 2000, 2, ""), (16384, 16384, 16384, 1, ""),
    (32768, 32768, 32768, 2, ""), (32768, 32768, 32768, 4, ""), (32768, 32768, 32768, 8, ""),
    (8192, 8192, 1048576, 8, ""), (8192, 8192, 1048576, 4, ""), (8192, 8192, 1048576, 2, ""),
    (16384, 16384, 16384, 8, ""), (16384, 16384, 16384, 8, "sk32,sm16,pk2,pm4"),
    (17408, 17408, 3473408, 128, ""), (1000, 3000, 5000, 6, ""), (1237, 4096, 777, 8, ""), (50000, 300, 300, 16, ""),
    (2001, 2003, 1999, 7, ""), (5000, 5000, 5000, 12, ""), (65536, 1024, 1024, 8, ""), (1024, 65536, 4096, 8, ""),
]

# automatic strategies under a memory limit (elements per rank) -> sequential steps get inserted
MEMORY_LIMITED_CASES = [
    (17408, 17408, 3473408, 4608, 52428800),   # tests/mapper.cpp:363-387 golden (RPA 128 water molecules)
    (4096, 4096, 4096, 8, 3 * 4096 * 4096 // 8 + 4096 * 512),
    (10000, 10000, 10000, 4, 80000000),
    (8192, 8192, 65536, 8, 200000000),
]
