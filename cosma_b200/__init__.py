"""cosma_b200 -- B200-native implementation of the COSMA distributed-GEMM hot path.

The product is libcosma_b200.so (C ABI, include/cosma_b200.h). This Python package is a thin host-side
mirror of the reference's interface used by the tests and bench; PyTorch is used only for device
memory, streams and torch.distributed plumbing."""
from ._lib import load, CosmaB200Error, LIB_PATH  # noqa: F401
