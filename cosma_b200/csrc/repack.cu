// Repack of operands the TMA path cannot address (see repack.h): one coalesced copy kernel per operand.
#include "repack.h"

namespace cosma_b200 {

namespace {
// column-major rows x cols, element = T (4, 8 or 16 bytes); every warp moves 32 consecutive elements of a column per step, so reads are
// coalesced whatever the source pitch and base alignment, and writes are coalesced and 16-byte aligned per column
template <typename T>
__global__ void repack_kernel(const T* __restrict__ src, int64_t ld_src, T* __restrict__ dst, int64_t ld_dst, int64_t rows, int64_t cols) {
    const int64_t row_chunks = (rows + blockDim.x - 1) / blockDim.x;
    const int64_t total = row_chunks * cols;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
        const int64_t c = w / row_chunks, r = (w % row_chunks) * blockDim.x + threadIdx.x;
        if (r < rows) dst[c * ld_dst + r] = src[c * ld_src + r];
    }
}

template <typename T>
cudaError_t launch(cudaStream_t stream, const void* src, int64_t ld, void* dst, int64_t ld_dst, int64_t rows, int64_t cols) {
    const int threads = 256;
    const int64_t work = ((rows + threads - 1) / threads) * cols;
    const int blocks = static_cast<int>(work < 148 * 16 ? (work > 0 ? work : 1) : 148 * 16);
    repack_kernel<T><<<blocks, threads, 0, stream>>>(static_cast<const T*>(src), ld, static_cast<T*>(dst), ld_dst, rows, cols);
    return cudaGetLastError();
}
}  // namespace

int repack_mode() {
    static const int mode = [] {
        const char* v = std::getenv("COSMA_B200_REPACK_UNALIGNED");
        if (!v || !*v) return 2;
        if (!std::strcmp(v, "ON") || !std::strcmp(v, "on") || !std::strcmp(v, "1") || !std::strcmp(v, "TRUE") || !std::strcmp(v, "true")) return 1;
        if (!std::strcmp(v, "AUTO") || !std::strcmp(v, "auto")) return 2;
        return 0;
    }();
    return mode;
}

cudaError_t repack_operand(cudaStream_t stream, const void* src, int64_t ld, int64_t rows, int64_t cols, int elem_bytes, int ld_multiple, Repacked& out) {
    out.ld = (rows + ld_multiple - 1) / ld_multiple * ld_multiple;
    if (out.ld < 1) out.ld = ld_multiple;
    const size_t bytes = static_cast<size_t>(out.ld) * static_cast<size_t>(cols > 0 ? cols : 1) * elem_bytes;
    cudaError_t e = cudaMallocAsync(&out.ptr, bytes, stream);
    if (e != cudaSuccess) { out.ptr = nullptr; return e; }
    if (rows > 0 && cols > 0) {
        // 16-byte elements need a 16-byte aligned source for the vector access: fall back to the 8-byte view (twice the rows)
        if (elem_bytes == 16 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) e = launch<double2>(stream, src, ld, out.ptr, out.ld, rows, cols);
        else if (elem_bytes == 16) e = launch<double>(stream, src, 2 * ld, out.ptr, 2 * out.ld, 2 * rows, cols);
        else if (elem_bytes == 8 && (reinterpret_cast<uintptr_t>(src) & 7) == 0) e = launch<double>(stream, src, ld, out.ptr, out.ld, rows, cols);
        else if (elem_bytes == 8) e = launch<float>(stream, src, 2 * ld, out.ptr, 2 * out.ld, 2 * rows, cols);
        else e = launch<float>(stream, src, ld, out.ptr, out.ld, rows, cols);
    }
    if (e != cudaSuccess) { cudaFreeAsync(out.ptr, stream); out.ptr = nullptr; }
    return e;
}

void repack_release(cudaStream_t stream, Repacked& r) {
    if (r.ptr) cudaFreeAsync(r.ptr, stream);
    r.ptr = nullptr;
}

}  // namespace cosma_b200
