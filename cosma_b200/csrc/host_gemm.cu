// Host-resident operand mode of the local GEMM: what a caller of the reference actually has.
// The reference's GPU path takes HOST pointers and streams <=5000^3 tiles through cuBLAS with 2 streams
// (libs/Tiled-MM/src/Tiled-MM/tiled_mm.cpp:270-365, 492-624). Here: A goes to HBM once, then column panels of
// B/C are pipelined -- H2D of panel j+1 and D2H of panel j-1 overlap the DMMA kernel on panel j -- so every
// operand byte crosses PCIe exactly once and the GEMM itself runs from HBM.
#include "gemm_f64_sm100.h"

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

int ensure_ws(size_t bytes, int n_events) {
    if (!g_ws.copy_in) {
        if (cudaStreamCreateWithFlags(&g_ws.copy_in, cudaStreamNonBlocking) != cudaSuccess) return COSMA_B200_CUDA_ERROR;
        if (cudaStreamCreateWithFlags(&g_ws.copy_out, cudaStreamNonBlocking) != cudaSuccess) return COSMA_B200_CUDA_ERROR;
    }
    if (g_ws.bytes < bytes) {
        if (g_ws.dev) cudaFree(g_ws.dev);
        g_ws.dev = nullptr;
        g_ws.bytes = 0;
        if (cudaMalloc(&g_ws.dev, bytes) != cudaSuccess) return COSMA_B200_OUT_OF_MEMORY;
        g_ws.bytes = bytes;
    }
    while (static_cast<int>(g_ws.ev.size()) < n_events) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return COSMA_B200_CUDA_ERROR;
        g_ws.ev.push_back(e);
    }
    return COSMA_B200_OK;
}

}  // namespace

void release_host_gemm_workspace() {
    if (g_ws.dev) cudaFree(g_ws.dev);
    g_ws.dev = nullptr;
    g_ws.bytes = 0;
}

// elem_doubles: 1 (double) or 2 (complex<double>). NN only (the reference base case is always 'N','N').
int gemm_f64_host(cudaStream_t stream, int elem_doubles, int64_t m, int64_t n, int64_t k, const double* alpha,
                  const double* A, int64_t lda, const double* B, int64_t ldb, const double* beta, double* C, int64_t ldc,
                  int* launches) {
    if (launches) *launches = 0;
    if (m < 0 || n < 0 || k < 0) return COSMA_B200_INVALID_ARG;
    if (m == 0 || n == 0) return COSMA_B200_OK;
    const size_t es = sizeof(double) * elem_doubles;
    const bool cplx = elem_doubles == 2;
    const bool beta_zero = beta[0] == 0.0 && (!cplx || beta[1] == 0.0);
    // device leading dimensions: compact and even (TMA needs 16-byte strides)
    const int64_t dlda = (m + 1) & ~int64_t(1), dldb = (std::max<int64_t>(k, 1) + 1) & ~int64_t(1), dldc = dlda;
    const int64_t panel = std::min<int64_t>(n, 2048);
    const int n_panels = static_cast<int>((n + panel - 1) / panel);
    const size_t a_bytes = size_t(dlda) * std::max<int64_t>(k, 1) * es;
    const size_t b_bytes = size_t(dldb) * n * es;
    const size_t c_bytes = size_t(dldc) * n * es;
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    int st = ensure_ws(up(a_bytes) + up(b_bytes) + up(c_bytes), 3 * n_panels + 2);
    if (st != COSMA_B200_OK) return st;
    char* base = static_cast<char*>(g_ws.dev);
    double* dA = reinterpret_cast<double*>(base);
    double* dB = reinterpret_cast<double*>(base + up(a_bytes));
    double* dC = reinterpret_cast<double*>(base + up(a_bytes) + up(b_bytes));
    cudaStream_t cin = g_ws.copy_in, cout = g_ws.copy_out;
    cudaEvent_t* ev = g_ws.ev.data();
    cudaEvent_t ev_start = ev[3 * n_panels], ev_a = ev[3 * n_panels + 1];

    // order the copy streams after whatever the caller queued on `stream`
    cudaEventRecord(ev_start, stream);
    cudaStreamWaitEvent(cin, ev_start, 0);
    cudaStreamWaitEvent(cout, ev_start, 0);
    if (k > 0) cudaMemcpy2DAsync(dA, dlda * es, A, lda * es, m * es, k, cudaMemcpyHostToDevice, cin);
    cudaEventRecord(ev_a, cin);
    cudaStreamWaitEvent(stream, ev_a, 0);
    for (int j = 0; j < n_panels; ++j) {
        const int64_t j0 = j * panel, w = std::min<int64_t>(panel, n - j0);
        if (k > 0)
            cudaMemcpy2DAsync(dB + j0 * dldb * elem_doubles, dldb * es, B + j0 * ldb * elem_doubles, ldb * es, k * es, w,
                              cudaMemcpyHostToDevice, cin);
        if (!beta_zero)
            cudaMemcpy2DAsync(dC + j0 * dldc * elem_doubles, dldc * es, C + j0 * ldc * elem_doubles, ldc * es, m * es, w,
                              cudaMemcpyHostToDevice, cin);
        cudaEventRecord(ev[3 * j], cin);
        cudaStreamWaitEvent(stream, ev[3 * j], 0);
        int path = 0;
        if (!cplx)
            st = dgemm_sm100(stream, 'N', 'N', m, w, k, alpha[0], dA, dlda, dB + j0 * dldb, dldb, beta[0], dC + j0 * dldc, dldc, &path);
        else
            st = zgemm_sm100(stream, 'N', 'N', m, w, k, alpha, dA, dlda, dB + 2 * j0 * dldb, dldb, beta, dC + 2 * j0 * dldc, dldc, &path);
        if (st != COSMA_B200_OK) return st;
        if (launches && path) ++*launches;
        cudaEventRecord(ev[3 * j + 1], stream);
        cudaStreamWaitEvent(cout, ev[3 * j + 1], 0);
        cudaMemcpy2DAsync(C + j0 * ldc * elem_doubles, ldc * es, dC + j0 * dldc * elem_doubles, dldc * es, m * es, w,
                          cudaMemcpyDeviceToHost, cout);
    }
    // the caller's stream completes only when the last panel is back on the host
    cudaEventRecord(ev[3 * (n_panels - 1) + 2], cout);
    cudaStreamWaitEvent(stream, ev[3 * (n_panels - 1) + 2], 0);
    return cudaGetLastError() == cudaSuccess ? COSMA_B200_OK : COSMA_B200_CUDA_ERROR;
}

}  // namespace cosma_b200
