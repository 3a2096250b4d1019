"""Generates tests/golden/planning_golden.json from the UNMODIFIED reference (oracle/_ref/libcosma_ref.so):
Strategy step strings, ranks used, memory_used and the complete Mapper layouts of A, B, C.
Run in the build container (needs /root/reference): python tests/golden/make_planning_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cases import BASELINE_CASES, MEMORY_LIMITED_CASES, REFERENCE_MULTIPLY_CASES, SCALAR_MATMUL_CASE  # noqa: E402
from oracle import oracle as orc  # noqa: E402

cases = [(m, n, k, P, s, 0) for (m, n, k, P, s) in REFERENCE_MULTIPLY_CASES + [SCALAR_MATMUL_CASE] + BASELINE_CASES]
cases += [(m, n, k, P, "", mem) for (m, n, k, P, mem) in MEMORY_LIMITED_CASES]
out = []
for m, n, k, P, steps, mem in cases:
    rec = {"m": m, "n": n, "k": k, "P": P, "steps_in": steps, "mem_limit": mem, "throws": False, "layout": {}}
    try:
        s, P_used, mem_used = orc.ref_strategy(m, n, k, P, mem, steps)
    except RuntimeError:
        rec["throws"] = True
        out.append(rec)
        continue
    rec.update({"steps": s, "P_used": P_used, "memory_used": mem_used})
    if P_used <= 16:
        for label in "ABC":
            rec["layout"][label] = [[list(b) for b in blocks] for blocks in orc.ref_mapper_layout(label, m, n, k, P_used, s)]
    out.append(rec)
with open(os.path.join(HERE, "planning_golden.json"), "w") as f:
    json.dump({"source": "eth-cscs/COSMA v2.8.4 @ /root/reference, built by oracle/Makefile", "cases": out}, f)
print("wrote", len(out), "cases")
