"""The programs bench.py runs at 32768^3 -- the overlapped micro-op programs of both transports and the serial schedule, at 2 / 4 / 8
ranks -- executed for all ranks in lock-step on the CPU with the m dimension compressed 2048 : 1 (tests/full_size_replay.py: the
programs never cut m, so the compression is an isomorphism; B keeps its full 32768 x 32768). Workspaces are poisoned with NaN, C holds
NaN when beta == 0, operands are integer-valued: the gathered result must equal the dense product EXACTLY. Together with the static
hazard analysis (tests/test_overlap_hazards_cpu.py) this is what stands behind the N = 4 line, whose program had not run on hardware at
bench size when the round's GPU budget ended (N = 2 and N = 8 passed the bench's own exact parity check on GPUs)."""
import numpy as np
import pytest

import full_size_replay as R

N = 32768


@pytest.fixture(scope="module")
def big_b():
    try:
        import psutil
        if psutil.virtual_memory().available < 24 * 2 ** 30:
            pytest.skip("needs about 20 GB of host memory")
    except ImportError:
        pass
    return R.random_b(N, N)


@pytest.mark.parametrize("P,transport,beta,alpha", [(2, "copy_engine", 0.0, 1.0), (4, "copy_engine", 0.0, 1.0), (8, "copy_engine", 0.0, 1.0),
                                                    (4, "copy_engine", -1.0, 2.0), (4, "nccl", 2.0, 1.0)])
def test_bench_programs_replayed(lib, monkeypatch, big_b, P, transport, beta, alpha):
    for v in ("COSMA_OVERLAP_COMM_AND_COMP", "COSMA_B200_OVERLAP_GRANULE", "COSMA_B200_OVERLAP_SMS", "COSMA_B200_OVERLAP_GBPS", "COSMA_B200_OVERLAP_ZERO_SM"):
        monkeypatch.delenv(v, raising=False)
    if transport == "copy_engine":
        monkeypatch.setenv("COSMA_B200_OVERLAP_ZERO_SM", "ON")  # the program cosma_b200_plan_bind_arenas switches to
    got, want, info = R.replay(N, N, N, P, alpha=alpha, beta=beta, overlapped=transport != "serial", Bg=big_b)
    assert info["strategy"] == {2: "pk2", 4: "pn2,pk2", 8: "pm2,pn2,pk2"}[P]
    if transport != "serial":
        assert all(g >= 3 for g in info["gemm_panels"]), info
    assert np.array_equal(got, want), (P, transport, beta, info)
