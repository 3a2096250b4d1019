// Host-resident operand mode of the local GEMM: what a caller of the reference actually has.
// The reference's GPU path takes HOST pointers and streams <=5000^3 tiles through cuBLAS with 2 streams
// (libs/Tiled-MM/src/Tiled-MM/tiled_mm.cpp:270-365, 492-624), re-sending A and B tiles for every (m, n, k) tile.
// Here every operand byte crosses PCIe exactly once and nearly all of it under the kernel:
//   * B (and C) move in column panels: H2D of panel j+1 and D2H of panel j-1 overlap the GEMM on panel j;
//   * A moves in ROW chunks while the first (wide) panel is computed chunk by chunk, C(rows_i, panel_0) =
//     A(rows_i, :) * B(:, panel_0), so the kernel starts after one B panel + one A chunk are on the device. The first
//     panel is wide enough that a chunk's GEMM lasts longer than its copy (the copy engine, not the SMs, waits);
//   * the last panel is thin so the D2H tail is short.
// Chunking along m and n only (never k) keeps every C element's summation order that of the one-launch GEMM:
// the result is bit-identical to the device-resident path.
#include "gemm_f64_sm100.h"
#include "gemm_tf32x3_sm100.h"
#include "host_stream.h"

#include <algorithm>
#include <mutex>

namespace cosma_b200 {
namespace {

struct HostGemmWorkspace {
    void* dev = nullptr;
    size_t bytes = 0;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    std::vector<cudaEvent_t> ev;
    ~HostGemmWorkspace() {}
};
thread_local HostGemmWorkspace g_ws;

int ensure_streams(int n_events) {
    if (!g_ws.copy_in) {
        if (cudaStreamCreateWithFlags(&g_ws.copy_in, cudaStreamNonBlocking) != cudaSuccess) return COSMA_B200_CUDA_ERROR;
        if (cudaStreamCreateWithFlags(&g_ws.copy_out, cudaStreamNonBlocking) != cudaSuccess) return COSMA_B200_CUDA_ERROR;
    }
    while (static_cast<int>(g_ws.ev.size()) < n_events) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return COSMA_B200_CUDA_ERROR;
        g_ws.ev.push_back(e);
    }
    return COSMA_B200_OK;
}

int ensure_ws(size_t bytes) {
    if (g_ws.bytes < bytes) {
        if (g_ws.dev) cudaFree(g_ws.dev);
        g_ws.dev = nullptr;
        g_ws.bytes = 0;
        if (cudaMalloc(&g_ws.dev, bytes) != cudaSuccess) return COSMA_B200_OUT_OF_MEMORY;
        g_ws.bytes = bytes;
    }
    return COSMA_B200_OK;
}

}  // namespace

void release_host_gemm_workspace() {
    if (g_ws.dev) cudaFree(g_ws.dev);
    g_ws.dev = nullptr;
    g_ws.bytes = 0;
}

// One 'N','N' GEMM of any of the four types with the scalars as doubles (the convention of cosma_b200_multiply).
int launch_gemm_nn(char dtype, cudaStream_t stream, int64_t m, int64_t n, int64_t k, const double* alpha, const void* A, int64_t lda,
                   const void* B, int64_t ldb, const double* beta, void* C, int64_t ldc, int* path) {
    switch (dtype) {
        case 'd':
            return dgemm_sm100(stream, 'N', 'N', m, n, k, alpha[0], static_cast<const double*>(A), lda, static_cast<const double*>(B), ldb,
                               beta[0], static_cast<double*>(C), ldc, path);
        case 'z':
            return zgemm_sm100(stream, 'N', 'N', m, n, k, alpha, static_cast<const double*>(A), lda, static_cast<const double*>(B), ldb, beta,
                               static_cast<double*>(C), ldc, path);
        case 's':
            return sgemm_sm100(stream, 'N', 'N', m, n, k, static_cast<float>(alpha[0]), static_cast<const float*>(A), lda,
                               static_cast<const float*>(B), ldb, static_cast<float>(beta[0]), static_cast<float*>(C), ldc, path);
        case 'c': {
            const float af[2] = {static_cast<float>(alpha[0]), static_cast<float>(alpha[1])};
            const float bf[2] = {static_cast<float>(beta[0]), static_cast<float>(beta[1])};
            return cgemm_sm100(stream, 'N', 'N', m, n, k, af, static_cast<const float*>(A), lda, static_cast<const float*>(B), ldb, bf,
                               static_cast<float*>(C), ldc, path);
        }
        default:
            return COSMA_B200_INVALID_ARG;
    }
}

// Column panels of C/B: {first column, width}. The first panel is wide when A is streamed under it (see the file
// header: width >= 4*R/BW elements keeps a row chunk's GEMM longer than its copy, R = kernel flop rate, BW = PCIe
// rate; ~2750 for FP64 at 37 TFLOP/s over ~54 GB/s), the last is thin when C goes back to the host.
std::vector<std::pair<int64_t, int64_t>> stream_panels(char dtype, int64_t n, bool a_streamed, bool c_to_host) {
    int64_t first = 2048;
    if (a_streamed) first = dtype == 'd' ? 3072 : (dtype == 'z' ? 1536 : (dtype == 's' ? 5632 : 3072));
    const int64_t mid = 2048, tail = 1024;
    std::vector<std::pair<int64_t, int64_t>> out;
    int64_t j = 0;
    bool is_first = true;
    while (j < n) {
        const int64_t rem = n - j;
        int64_t w = std::min(rem, is_first ? first : mid);
        if (c_to_host && w == rem && rem > tail + tail / 2) w = rem - tail;  // split the end: (rem - 1024, 1024)
        out.emplace_back(j, w);
        j += w;
        is_first = false;
    }
    return out;
}

int stream_gemm(cudaStream_t stream, const StreamGemmArgs& g, int* launches) {
    if (launches) *launches = 0;
    if (g.m < 0 || g.n < 0 || g.k < 0) return COSMA_B200_INVALID_ARG;
    if (g.m == 0 || g.n == 0) return COSMA_B200_OK;
    const bool cplx = g.dtype == 'z' || g.dtype == 'c';
    const size_t es = (g.dtype == 'd' || g.dtype == 'z' ? 8 : 4) * (cplx ? 2 : 1);
    const bool beta_zero = g.beta[0] == 0.0 && (!cplx || g.beta[1] == 0.0);
    const bool a_in = g.hA && g.k > 0, b_in = g.hB && g.k > 0, c_in = g.hC_in && !beta_zero, c_out = g.hC_out != nullptr;
    int st, path = 0;
    if (!a_in && !b_in && !c_in && !c_out) {  // nothing crosses PCIe: one launch
        st = launch_gemm_nn(g.dtype, stream, g.m, g.n, g.k, g.alpha, g.dA, g.dlda, g.dB, g.dldb, g.beta, g.dC, g.dldc, &path);
        if (launches && path) ++*launches;
        return st;
    }
    const auto panels = stream_panels(g.dtype, g.n, a_in, c_out);
    const int n_panels = static_cast<int>(panels.size());
    const int64_t row_chunk = 768;  // 6 tile rows x 24 tile columns of the first panel = 144 tiles ~ one wave of 148 CTAs
    const int n_chunks = a_in ? static_cast<int>((g.m + row_chunk - 1) / row_chunk) : 0;
    st = ensure_streams(2 * n_panels + n_chunks + 2);
    if (st != COSMA_B200_OK) return st;
    cudaStream_t cin = g_ws.copy_in, cout = g_ws.copy_out;
    cudaEvent_t* ev_in = g_ws.ev.data();                  // [n_panels] panel operands on the device
    cudaEvent_t* ev_done = ev_in + n_panels;              // [n_panels] panel computed
    cudaEvent_t* ev_chunk = ev_done + n_panels;           // [n_chunks] A row chunk on the device
    cudaEvent_t ev_start = ev_chunk[n_chunks], ev_end = ev_chunk[n_chunks + 1];
    auto at = [es](const void* p, int64_t elems) { return static_cast<const char*>(p) + elems * es; };
    auto atw = [es](void* p, int64_t elems) { return static_cast<char*>(p) + elems * es; };

    // order the copy streams after whatever the caller queued on `stream` (previous users of the device buffers)
    cudaEventRecord(ev_start, stream);
    cudaStreamWaitEvent(cin, ev_start, 0);
    if (c_out) cudaStreamWaitEvent(cout, ev_start, 0);
    for (int j = 0; j < n_panels; ++j) {
        const int64_t j0 = panels[j].first, w = panels[j].second;
        if (b_in) cudaMemcpy2DAsync(atw(g.dB, j0 * g.dldb), g.dldb * es, at(g.hB, j0 * g.ldb), g.ldb * es, g.k * es, w, cudaMemcpyHostToDevice, cin);
        if (c_in) cudaMemcpy2DAsync(atw(g.dC, j0 * g.dldc), g.dldc * es, at(g.hC_in, j0 * g.ldc), g.ldc * es, g.m * es, w, cudaMemcpyHostToDevice, cin);
        if (b_in || c_in) {
            cudaEventRecord(ev_in[j], cin);
            cudaStreamWaitEvent(stream, ev_in[j], 0);
        }
        if (j == 0 && a_in) {
            for (int i = 0; i < n_chunks; ++i) {
                const int64_t i0 = i * row_chunk, h = std::min(row_chunk, g.m - i0);
                cudaMemcpy2DAsync(atw(g.dA, i0), g.dlda * es, at(g.hA, i0), g.lda * es, h * es, g.k, cudaMemcpyHostToDevice, cin);
                cudaEventRecord(ev_chunk[i], cin);
                cudaStreamWaitEvent(stream, ev_chunk[i], 0);
                st = launch_gemm_nn(g.dtype, stream, h, w, g.k, g.alpha, at(g.dA, i0), g.dlda, at(g.dB, j0 * g.dldb), g.dldb, g.beta,
                                    atw(g.dC, j0 * g.dldc + i0), g.dldc, &path);
                if (st != COSMA_B200_OK) return st;
                if (launches && path) ++*launches;
            }
        } else {
            st = launch_gemm_nn(g.dtype, stream, g.m, w, g.k, g.alpha, g.dA, g.dlda, at(g.dB, j0 * g.dldb), g.dldb, g.beta,
                                atw(g.dC, j0 * g.dldc), g.dldc, &path);
            if (st != COSMA_B200_OK) return st;
            if (launches && path) ++*launches;
        }
        if (c_out) {
            cudaEventRecord(ev_done[j], stream);
            cudaStreamWaitEvent(cout, ev_done[j], 0);
            cudaMemcpy2DAsync(atw(g.hC_out, j0 * g.ldc), g.ldc * es, at(g.dC, j0 * g.dldc), g.dldc * es, g.m * es, w, cudaMemcpyDeviceToHost, cout);
        }
    }
    if (c_out) {  // the caller's stream completes only when the last panel is back on the host
        cudaEventRecord(ev_end, cout);
        cudaStreamWaitEvent(stream, ev_end, 0);
    }
    return cudaGetLastError() == cudaSuccess ? COSMA_B200_OK : COSMA_B200_CUDA_ERROR;
}

// elem_doubles: 1 (double) or 2 (complex<double>). NN only (the reference base case is always 'N','N').
int gemm_f64_host(cudaStream_t stream, int elem_doubles, int64_t m, int64_t n, int64_t k, const double* alpha,
                  const double* A, int64_t lda, const double* B, int64_t ldb, const double* beta, double* C, int64_t ldc,
                  int* launches) {
    if (launches) *launches = 0;
    if (m < 0 || n < 0 || k < 0) return COSMA_B200_INVALID_ARG;
    if (m == 0 || n == 0) return COSMA_B200_OK;
    const size_t es = sizeof(double) * elem_doubles;
    // device leading dimensions: compact and even (TMA needs 16-byte strides)
    const int64_t dlda = (m + 1) & ~int64_t(1), dldb = (std::max<int64_t>(k, 1) + 1) & ~int64_t(1), dldc = dlda;
    const size_t a_bytes = size_t(dlda) * std::max<int64_t>(k, 1) * es;
    const size_t b_bytes = size_t(dldb) * n * es;
    const size_t c_bytes = size_t(dldc) * n * es;
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    int st = ensure_ws(up(a_bytes) + up(b_bytes) + up(c_bytes));
    if (st != COSMA_B200_OK) return st;
    char* base = static_cast<char*>(g_ws.dev);
    StreamGemmArgs g;
    g.dtype = elem_doubles == 2 ? 'z' : 'd';
    g.m = m; g.n = n; g.k = k;
    g.alpha = alpha; g.beta = beta;
    g.dA = base; g.dlda = dlda;
    g.dB = base + up(a_bytes); g.dldb = dldb;
    g.dC = base + up(a_bytes) + up(b_bytes); g.dldc = dldc;
    g.hA = A; g.lda = lda;
    g.hB = B; g.ldb = ldb;
    g.hC_in = C; g.hC_out = C; g.ldc = ldc;
    return stream_gemm(stream, g, launches);
}

}  // namespace cosma_b200
