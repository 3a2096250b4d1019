// Operands the TMA path cannot address (base not 16-byte aligned, or a row pitch that is not a multiple of 16 bytes -- COSMA's native
// layout has ld = local rows, so irregular splits produce odd ones) are copied ONCE into a stream-ordered scratch with a legal
// pitch by a coalesced copy kernel (repack.cu, HBM-bound); the tensor-pipe kernel then runs on the copy.
// COSMA_B200_REPACK_UNALIGNED = AUTO (default) | ON | OFF. Measured on B200 (profiles/r2_repack_probe.jsonl, 8191 x 8192 x 8191 FP64): the
// generic kernel reaches 12.8 TFLOP/s against 36 for the tensor-pipe kernel, so two extra passes over A and B (16 (mk + kn) bytes at
// ~4 TB/s) pay as soon as the product is more than a few hundred cubed: AUTO repacks when m n k >= 2^27 and k >= 64; ON whenever the
// kernel's tiles are not mostly padding (k >= 64, m n >= 128^2); OFF: the generic kernel, as in round 1.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace cosma_b200 {

int repack_mode();  // 0 off, 1 on, 2 auto

inline bool repack_wanted(int64_t m, int64_t n, int64_t k) {
    const int mode = repack_mode();
    if (mode == 0 || k < 64 || m * n < 128 * 128) return false;
    return mode == 1 || static_cast<double>(m) * static_cast<double>(n) * static_cast<double>(k) >= 134217728.0;
}

struct Repacked {
    void* ptr = nullptr;   // scratch (cudaMallocAsync on the GEMM's stream), nullptr when the operand was fine as it was
    int64_t ld = 0;        // leading dimension of the scratch, in elements
};

// Copies the stored rows x cols operand (column-major, leading dimension ld, elements of elem_bytes) into a scratch whose leading
// dimension is rows rounded up to a multiple of ld_multiple elements. Returns cudaSuccess and out.ptr != nullptr on success.
cudaError_t repack_operand(cudaStream_t stream, const void* src, int64_t ld, int64_t rows, int64_t cols, int elem_bytes, int ld_multiple,
                           Repacked& out);
void repack_release(cudaStream_t stream, Repacked& r);

}  // namespace cosma_b200
