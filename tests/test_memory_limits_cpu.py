"""Memory-limited automatic strategies. (1) COSMA_CPU_MAX_MEMORY keeps the reference's meaning inside the plan: the
strategy equals the reference's Strategy under the same limit. (2) COSMA_B200_DEVICE_MEMORY_MB / cosma_b200_fit_strategy: sequential
steps are added until the arenas of the COMPILED schedule -- what is really allocated in HBM -- fit the budget on every rank, and the
resulting schedule still multiplies correctly (lock-step simulation)."""
import ctypes
import os

import numpy as np
import pytest

import schedule_sim
from cosma_b200 import planning
from cosma_b200.distributed import MultiplyPlan


def _fit(lib, m, n, k, P, eb, budget_bytes, prefix=""):
    out = ctypes.create_string_buffer(4096)
    Pout, foot = ctypes.c_int(0), ctypes.c_longlong(0)
    rc = lib.cosma_b200_fit_strategy(m, n, k, P, prefix.encode(), eb, ctypes.c_longlong(budget_bytes), out, 4096, ctypes.byref(Pout), ctypes.byref(foot))
    return rc, out.value.decode(), Pout.value, foot.value


def _footprint_bytes(m, n, k, P, steps, dtype="d"):
    eb = {"d": 8, "z": 16, "s": 4, "c": 8}[dtype]
    worst = 0
    for r in range(P):
        pl = MultiplyPlan(None, m, n, k, steps, dtype, rank=r, nranks=P, allocate=False)
        worst = max(worst, sum(pl.arena_elements) * eb)
        pl.destroy()
    return worst


@pytest.mark.parametrize("m,n,k,P,mb", [(8192, 8192, 8192, 8, 400), (4096, 4096, 65536, 4, 1100), (10000, 10000, 10000, 4, 1200)])
def test_reference_memory_switch_inside_the_plan(lib, ref, m, n, k, P, mb, monkeypatch):
    R = ref.ref()
    monkeypatch.setenv("COSMA_CPU_MAX_MEMORY", str(mb))
    pl = MultiplyPlan(None, m, n, k, "", "d", rank=0, nranks=P, allocate=False)
    limited = pl.strategy
    pl.destroy()
    monkeypatch.delenv("COSMA_CPU_MAX_MEMORY")
    pl = MultiplyPlan(None, m, n, k, "", "d", rank=0, nranks=P, allocate=False)
    free = pl.strategy
    pl.destroy()
    limit_elems = mb * 1024 * 1024 // 8
    out = ctypes.create_string_buffer(4096)
    Pout, mem = ctypes.c_int(0), ctypes.c_longlong(0)
    assert R.ref_strategy(m, n, k, P, ctypes.c_longlong(limit_elems), b"", out, 4096, ctypes.byref(Pout), ctypes.byref(mem)) >= 0
    assert limited == out.value.decode()
    assert limited != free and "s" in limited.replace(",", "")[::1] and any(step.startswith("s") for step in limited.split(","))


@pytest.mark.parametrize("m,n,k,P,dtype", [(8192, 8192, 8192, 8, "d"), (16384, 4096, 4096, 4, "z"), (4096, 4096, 131072, 8, "d"), (6000, 7000, 5000, 6, "s")])
def test_fit_to_device_memory(lib, m, n, k, P, dtype):
    eb = {"d": 8, "z": 16, "s": 4, "c": 8}[dtype]
    free_steps = planning.strategy(m, n, k, P)[0]
    unlimited = _footprint_bytes(m, n, k, P, free_steps, dtype)
    # a budget that already fits changes nothing
    rc, steps, P_used, foot = _fit(lib, m, n, k, P, eb, unlimited)
    assert rc == 0 and steps == free_steps and foot == unlimited
    # half way between the local matrices alone and the unlimited footprint: sequential steps appear and every rank's arenas fit
    local_only = (m * k + k * n + m * n) * eb // P
    budget = (local_only + unlimited) // 2
    rc, steps, P_used, foot = _fit(lib, m, n, k, P, eb, budget)
    assert rc == 0 and any(s.startswith("s") for s in steps.split(",")), steps
    assert foot <= budget and _footprint_bytes(m, n, k, P, steps, dtype) == foot
    # the local matrices alone do not fit: a clear error, no strategy
    rc, _, _, _ = _fit(lib, m, n, k, P, eb, (m * k + k * n + m * n) * eb // P // 2)
    assert rc != 0 and "does not fit" in lib.cosma_b200_last_error().decode()


def test_device_memory_switch_inside_the_plan_and_still_correct(lib, monkeypatch):
    m, n, k, P = 1024, 768, 1280, 4
    free_steps = planning.strategy(m, n, k, P)[0]
    unlimited = _footprint_bytes(m, n, k, P, free_steps)
    budget_mb = max(1, ((m * k + k * n + m * n) * 8 // P + unlimited) // 2 // (1024 * 1024))
    monkeypatch.setenv("COSMA_B200_DEVICE_MEMORY_MB", str(budget_mb))
    pl = MultiplyPlan(None, m, n, k, "", "d", rank=0, nranks=P, allocate=False)
    steps = pl.strategy
    pl.destroy()
    monkeypatch.delenv("COSMA_B200_DEVICE_MEMORY_MB")
    assert steps != free_steps and _footprint_bytes(m, n, k, P, steps) <= budget_mb * 1024 * 1024
    got, want, P_used = schedule_sim.simulate(m, n, k, P, steps, alpha=1.0, beta=1.0)
    assert P_used == P and np.array_equal(got, want)
