"""The exact parity check bench.py runs after its timed loop (cosma_b200.distributed.MultiplyJob.parity), exercised on the CPU with a
stand-in plan whose multiply is a dense torch matmul: the check must accept the correct product and reject a single wrong element --
both through the sampled dot products (if the element is sampled) and, always, through the checksum of checksums."""
import pytest

torch = pytest.importorskip("torch")

from cosma_b200 import distributed  # noqa: E402


class _Mat:
    def __init__(self, n, dtype):
        self.local = torch.zeros(n, dtype=dtype)
        self.initial = n


class _Plan:
    """One rank owning everything (P = 1), or a rank owning a sub-block of C with the matching row / column panels."""

    def __init__(self, m, n, k, dtype, rows=None, cols=None, corrupt=None):
        self.m, self.n, self.k, self.idle = m, n, k, False
        self.rows, self.cols = rows or (0, m - 1), cols or (0, n - 1)
        nr, nc = self.rows[1] - self.rows[0] + 1, self.cols[1] - self.cols[0] + 1
        self.A, self.B, self.C = _Mat(nr * k, dtype), _Mat(k * nc, dtype), _Mat(nr * nc, dtype)
        self.nr, self.nc, self.corrupt = nr, nc, corrupt

    def local_blocks(self, label, rank=None):
        return {"A": [(self.rows[0], self.rows[1], 0, self.k - 1)], "B": [(0, self.k - 1, self.cols[0], self.cols[1])],
                "C": [(self.rows[0], self.rows[1], self.cols[0], self.cols[1])]}[label]

    def multiply(self, alpha, beta):
        a = self.A.local.reshape(self.k, self.nr).T
        b = self.B.local.reshape(self.nc, self.k).T
        c = (a @ b).T.reshape(-1).clone()
        if self.corrupt is not None:
            c[self.corrupt] += 1
        self.C.local.copy_(c)

    def overlap(self):
        return {"enabled": False, "why": "stand-in"}


def _job(plan, dtype):
    job = distributed.MultiplyJob.__new__(distributed.MultiplyJob)
    job.plan, job.device, job.rank, job.world = plan, torch.device("cpu"), 0, 1
    job.cplx = dtype in (torch.complex128, torch.complex64)
    return job


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128, torch.float32])
def test_parity_check_accepts_the_product_and_rejects_one_wrong_element(dtype):
    for rows, cols in ((None, None), ((32, 95), (64, 127))):
        good = _job(_Plan(128, 160, 96, dtype, rows, cols), dtype).parity(samples=8)
        assert good["ok"] and good["sampled_elements"] == 8, good
        bad = _job(_Plan(128, 160, 96, dtype, rows, cols, corrupt=1234), dtype).parity(samples=8)
        assert not bad["ok"]
