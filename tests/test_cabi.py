"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol include/cosma_b200.h
declares (no compute calls here)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cosma_b200.h")).read()
    return sorted(set(re.findall(r"COSMA_B200_API[^;(]*?\b(cosma_b200_\w+)\s*\(", text)))


def test_header_declares_something():
    syms = declared_symbols()
    assert "cosma_b200_dgemm" in syms and "cosma_b200_version" in syms


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, "declared in include/cosma_b200.h but not exported: %s" % missing


def test_version_and_error_strings(lib):
    assert lib.cosma_b200_version().decode().startswith("cosma_b200")
    assert lib.cosma_b200_last_error() is not None


def test_no_oracle_in_product():
    """The product path must never route through the oracle or a CPU fallback."""
    bad = []
    for root, _, files in os.walk(os.path.join(ROOT, "cosma_b200")):
        if os.sep + "build" in root:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", ".cuh")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M) or "liboracle" in txt or "libcosma_ref" in txt:
                    bad.append(os.path.join(root, f))
    assert not bad, bad
