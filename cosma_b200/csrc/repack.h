// Operands the TMA path cannot address (base not 16-byte aligned, or a row pitch that is not a multiple of 16 bytes -- COSMA's native
// layout has ld = local rows, so irregular splits produce odd ones) are copied ONCE into a stream-ordered scratch with a legal
// pitch; the tensor-pipe kernel then runs on the copy. One HBM-bound 2-D device copy per operand against an O(mnk) GEMM.
// Opt-in (COSMA_B200_REPACK_UNALIGNED=ON) until measured on a GPU; off: the generic kernel, as before.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace cosma_b200 {

inline bool repack_unaligned_enabled() {
    static const bool on = [] {
        const char* v = std::getenv("COSMA_B200_REPACK_UNALIGNED");
        return v && (!std::strcmp(v, "ON") || !std::strcmp(v, "on") || !std::strcmp(v, "1") || !std::strcmp(v, "TRUE") || !std::strcmp(v, "true"));
    }();
    return on;
}

struct Repacked {
    void* ptr = nullptr;   // scratch (cudaMallocAsync on the GEMM's stream), nullptr when the operand was fine as it was
    int64_t ld = 0;        // leading dimension of the scratch, in elements
};

// Copies the stored rows x cols operand (column-major, leading dimension ld, elements of elem_bytes) into a scratch whose leading
// dimension is rows rounded up to a multiple of ld_multiple elements. Returns cudaSuccess and out.ptr != nullptr on success.
inline cudaError_t repack_operand(cudaStream_t stream, const void* src, int64_t ld, int64_t rows, int64_t cols, int elem_bytes, int ld_multiple,
                                  Repacked& out) {
    out.ld = (rows + ld_multiple - 1) / ld_multiple * ld_multiple;
    if (out.ld < 1) out.ld = ld_multiple;
    const size_t bytes = static_cast<size_t>(out.ld) * static_cast<size_t>(cols > 0 ? cols : 1) * elem_bytes;
    cudaError_t e = cudaMallocAsync(&out.ptr, bytes, stream);
    if (e != cudaSuccess) { out.ptr = nullptr; return e; }
    if (rows > 0 && cols > 0)
        e = cudaMemcpy2DAsync(out.ptr, static_cast<size_t>(out.ld) * elem_bytes, src, static_cast<size_t>(ld) * elem_bytes,
                              static_cast<size_t>(rows) * elem_bytes, static_cast<size_t>(cols), cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) { cudaFreeAsync(out.ptr, stream); out.ptr = nullptr; }
    return e;
}

inline void repack_release(cudaStream_t stream, Repacked& r) {
    if (r.ptr) cudaFreeAsync(r.ptr, stream);
    r.ptr = nullptr;
}

}  // namespace cosma_b200
