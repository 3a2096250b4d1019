// Executor of a cosma::Schedule on one GPU: the run-time half of cosma::multiply (reference
// src/cosma/multiply.cpp:243-314). Communication = NCCL over NVLink on per-step ring communicators
// (the reference's communicator::create_communicators, communicator.cpp:282-308, and gpu::nccl_copy / nccl_reduce,
// gpu/nccl_utils.cpp:45-280, which stage through host memory around every collective); local compute = the sm_100a
// DGEMM/ZGEMM kernels. Operands never leave HBM.
#include "exec_internal.h"

#include <cosma/auto_strategy.hpp>
#include "gemm_f64_sm100.h"
#include "gemm_tf32x3_sm100.h"
#include "host_stream.h"
#include "cta_budget.h"
#include "peer_transport.h"

#include <cstring>

namespace cosma_b200 {

namespace {

// C[i] = beta * C[i] + T[i]  (E1: the reference's host loop, two_sided_communicator.cpp:219-224)
template <typename R, bool CPLX>
__global__ void axpby_kernel(int64_t n, R br, R bi, R* __restrict__ C, const R* __restrict__ T) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    if (!CPLX) {
        for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += stride) C[i] = br * C[i] + T[i];
    } else {
        for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += stride) {
            const R cr = C[2 * i], ci = C[2 * i + 1];
            C[2 * i] = br * cr - bi * ci + T[2 * i];
            C[2 * i + 1] = br * ci + bi * cr + T[2 * i + 1];
        }
    }
}

#define NCCL_TRY COSMA_B200_NCCL_TRY
#define CUDA_TRY COSMA_B200_CUDA_TRY

// C[i] = (br, bi) * C[i] + T[i] over n elements of the plan's type
int launch_axpby(const Plan& plan, int64_t n, double br, double bi, char* C, const char* T, cudaStream_t stream) {
    if (n <= 0) return COSMA_B200_OK;
    const int threads = 256;
    int blocks = static_cast<int>(std::min<int64_t>((n + threads - 1) / threads, 148 * 8));
    if (blocks < 1) blocks = 1;
    const int E = plan.elem_reals;
    if (plan.real_bytes == 8) {
        if (E == 1) axpby_kernel<double, false><<<blocks, threads, 0, stream>>>(n, br, bi, reinterpret_cast<double*>(C), reinterpret_cast<const double*>(T));
        else axpby_kernel<double, true><<<blocks, threads, 0, stream>>>(n, br, bi, reinterpret_cast<double*>(C), reinterpret_cast<const double*>(T));
    } else {
        const float fr = static_cast<float>(br), fi = static_cast<float>(bi);
        if (E == 1) axpby_kernel<float, false><<<blocks, threads, 0, stream>>>(n, fr, fi, reinterpret_cast<float*>(C), reinterpret_cast<const float*>(T));
        else axpby_kernel<float, true><<<blocks, threads, 0, stream>>>(n, fr, fi, reinterpret_cast<float*>(C), reinterpret_cast<const float*>(T));
    }
    CUDA_TRY(cudaGetLastError());
    return COSMA_B200_OK;
}

int run_allgather(const Plan& plan, const cosma::ScheduleOp& op, char* arena, cudaStream_t stream) {
    const NcclApi* N = nccl();
    const int E = plan.elem_reals;                 // NCCL counts are in real scalars (complex = 2 x real, nccl_utils.cpp:105-107)
    const int64_t EB = plan.elem_bytes();
    const ncclDataType_t ndt = plan.real_bytes == 8 ? ncclDouble : ncclFloat;
    ncclComm_t comm = plan.ring_comms[op.ring_index];
    const int div = static_cast<int>(op.ring.size());
    const size_t nb = op.piece[0].size();
    const char* src = arena + op.src_off * EB;
    char* dst = arena + op.dst_off * EB;
    if (op.regular) {
        NCCL_TRY(N->AllGather(src, dst, static_cast<size_t>(op.piece[0][0]) * E, ndt, comm, stream));
        return COSMA_B200_OK;
    }
    // exact placement: bucket-major destination, member-major inside a bucket
    std::vector<int64_t> src_off(div, 0);
    int64_t dst_off = 0;
    NCCL_TRY(N->GroupStart());
    for (size_t b = 0; b < nb; ++b) {
        for (int g = 0; g < div; ++g) {
            const int64_t cnt = op.piece[g][b];
            if (cnt > 0) {
                if (g == op.my_pos) {
                    CUDA_TRY(cudaMemcpyAsync(dst + dst_off * EB, src + src_off[g] * EB, cnt * EB, cudaMemcpyDeviceToDevice, stream));
                } else {
                    NCCL_TRY(N->Recv(dst + dst_off * EB, static_cast<size_t>(cnt) * E, ndt, g, comm, stream));
                }
            }
            src_off[g] += cnt;
            dst_off += cnt;
        }
        // my piece of this bucket goes to every other member
        const int64_t mine = op.piece[op.my_pos][b];
        if (mine > 0)
            for (int g = 0; g < div; ++g)
                if (g != op.my_pos)
                    NCCL_TRY(N->Send(src + (src_off[op.my_pos] - mine) * EB, static_cast<size_t>(mine) * E, ndt, g, comm, stream));
    }
    NCCL_TRY(N->GroupEnd());
    return COSMA_B200_OK;
}

int run_reduce(const Plan& plan, const cosma::ScheduleOp& op, char* arena, const double* beta_user, cudaStream_t stream) {
    const NcclApi* N = nccl();
    const int E = plan.elem_reals;
    const int64_t EB = plan.elem_bytes();
    const ncclDataType_t ndt = plan.real_bytes == 8 ? ncclDouble : ncclFloat;
    ncclComm_t comm = plan.ring_comms[op.ring_index];
    const int div = static_cast<int>(op.ring.size());
    const size_t nb = op.piece[0].size();
    double br = 0.0, bi = 0.0;
    if (op.beta == cosma::BetaMode::ONE) br = 1.0;
    else if (op.beta == cosma::BetaMode::USER) { br = beta_user[0]; bi = E == 2 ? beta_user[1] : 0.0; }
    const bool beta_zero = br == 0.0 && bi == 0.0;
    const char* src = arena + op.src_off * EB;
    char* dst = arena + op.dst_off * EB;
    char* recv = beta_zero ? dst : arena + op.tmp_off * EB;
    int64_t mine_total = 0;
    for (auto v : op.piece[op.my_pos]) mine_total += v;
    if (op.regular) {
        NCCL_TRY(N->ReduceScatter(src, recv, static_cast<size_t>(op.piece[0][0]) * E, ndt, ncclSum, comm, stream));
    } else {
        // reduce-scatter-v with exact counts: one rooted reduction per (bucket, member), fused in one NCCL group
        int64_t off = 0, roff = 0;
        NCCL_TRY(N->GroupStart());
        for (size_t b = 0; b < nb; ++b)
            for (int g = 0; g < div; ++g) {
                const int64_t cnt = op.piece[g][b];
                if (cnt > 0)
                    NCCL_TRY(N->Reduce(src + off * EB, g == op.my_pos ? recv + roff * EB : nullptr, static_cast<size_t>(cnt) * E, ndt, ncclSum,
                                       g, comm, stream));
                if (g == op.my_pos) roff += cnt;
                off += cnt;
            }
        NCCL_TRY(N->GroupEnd());
    }
    if (!beta_zero && mine_total > 0) {
        const int st = launch_axpby(plan, mine_total, br, bi, dst, recv, stream);
        if (st != COSMA_B200_OK) return st;
    }
    return COSMA_B200_OK;
}

void beta_of(cosma::BetaMode mode, const double* user, int E, double out[2]) {
    out[0] = out[1] = 0.0;
    if (mode == cosma::BetaMode::ONE) out[0] = 1.0;
    else if (mode == cosma::BetaMode::USER) { out[0] = user[0]; out[1] = E == 2 ? user[1] : 0.0; }
}

// The overlapped form of a schedule (cosma::OverlapProgram, host/overlap.cpp): micro-ops issued on two streams -- the caller's for the
// GEMM panels, the plan's own high-priority stream for the NCCL kernels -- tied together by events. Narrow GEMMs leave plan.reserved
// SMs to the communication kernels, whose communicators are limited to as many CTAs, so neither side ever waits for the other's SMs.
int plan_run_overlapped(Plan& plan, const double* alpha, const double* beta, char* arenas[3], cudaStream_t stream) {
    const auto& prog = plan.overlap.ops;
    const auto& ops = plan.schedule.ops();
    const NcclApi* N = nccl();
    const int E = plan.elem_reals;
    const int64_t EB = plan.elem_bytes();
    const ncclDataType_t ndt = plan.real_bytes == 8 ? ncclDouble : ncclFloat;
    if (!plan.comm_stream) {
        int lo = 0, hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&plan.comm_stream, cudaStreamNonBlocking, hi));
    }
    while (plan.micro_ev.size() < 2 * prog.size() + 1) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        plan.micro_ev.push_back(e);
    }
    cudaStream_t comm = plan.comm_stream;
    cudaEvent_t entry = plan.micro_ev[2 * prog.size()];
    // the communication stream starts after whatever the caller queued (the operands are where they should be)
    CUDA_TRY(cudaEventRecord(entry, stream));
    CUDA_TRY(cudaStreamWaitEvent(comm, entry, 0));
    // zero-SM transport (peer_transport.h): copy-engine pushes into the ring mates' arenas instead of NCCL kernels
    PeerTransport* found = nullptr;
    for (auto& t : plan.peers)
        if (t->ready && arenas[0] == t->bound[0] && arenas[1] == t->bound[1] && arenas[2] == t->bound[2]) found = t.get();
    static PeerTransport none;
    PeerTransport& peer = found ? *found : none;
    const bool ce = found != nullptr;  // arenas that were never bound: NCCL kernels (every rank decides alike: the calls are collective)
    if (ce) {
        ++peer.epoch;
        for (const auto& link : peer.links) {  // every mate learns that this rank's previous call has drained
            const int st = peer_signal_entered(peer, link, comm);
            if (st != COSMA_B200_OK) return st;
        }
    }
    if (ce) {
        // Own pieces into their slots of the expanded buffers FIRST, while the device is still empty: an intra-device copy may run as a
        // kernel, and behind a persistent GEMM that holds every SM it would only start when that GEMM ends (measured at 8 GPUs,
        // profiles/r2d_bench_n8_default.json: 1.9 ms of the step exposed with the copies queued after the pushes).
        for (size_t i = 0; i < prog.size(); ++i) {
            const cosma::MicroOp& o = prog[i];
            if (o.stream != 1 || o.kind != cosma::MicroKind::ALLGATHER) continue;
            const auto& op = ops[o.op];
            const int64_t cnt = op.piece[0][0];
            CUDA_TRY(cudaMemcpyAsync(arenas[op.matrix] + (op.dst_off + op.my_pos * cnt) * EB, arenas[op.matrix] + op.src_off * EB,
                                     static_cast<size_t>(cnt * EB), cudaMemcpyDeviceToDevice, comm));
        }
    }
    std::vector<int> pending;  // allgathers whose pushes are queued but whose arrival has not been waited for yet
    int last_comm = -1;
    for (size_t i = 0; i < prog.size(); ++i) {
        const cosma::MicroOp& o = prog[i];
        cudaStream_t s = o.stream ? comm : stream;
        for (int w : o.wait)
            if (prog[w].stream != o.stream) CUDA_TRY(cudaStreamWaitEvent(s, plan.micro_ev[2 * w + 1], 0));
        if (plan.time_gemms) CUDA_TRY(cudaEventRecord(plan.micro_ev[2 * i], s));
        int st = COSMA_B200_OK;
        bool record_end = true;
        if (ce && o.stream == 1 && (o.kind == cosma::MicroKind::ALLGATHER || o.kind == cosma::MicroKind::EXCHANGE)) {
            const PeerLink* link = peer_link(peer, static_cast<int>(i));
            if (!link) return COSMA_B200_INTERNAL_ERROR;
            if (o.kind == cosma::MicroKind::ALLGATHER) {
                // own piece: to the mate (copy engine over NVLink) and into the own slot of the expanded buffer; every push is queued
                // before the first wait for an arrival, so the transfers of consecutive allgathers are all in flight together
                const auto& op = ops[o.op];
                const int64_t cnt = op.piece[0][0];
                const char* src = arenas[op.matrix] + op.src_off * EB;
                st = peer_push(peer, *link, src, static_cast<size_t>(cnt * EB), false, s);
                if (st != COSMA_B200_OK) return st;
                pending.push_back(static_cast<int>(i));
                const bool more = i + 1 < prog.size() && prog[i + 1].stream == 1 && prog[i + 1].kind == cosma::MicroKind::ALLGATHER;
                if (!more) {
                    for (int j : pending) {
                        st = peer_wait_arrived(peer, *peer_link(peer, j), s);
                        if (st != COSMA_B200_OK) return st;
                        CUDA_TRY(cudaEventRecord(plan.micro_ev[2 * j + 1], s));
                    }
                    pending.clear();
                }
                record_end = false;
            } else {
                double b[2];
                beta_of(o.beta, beta, E, b);
                st = peer_push(peer, *link, arenas[2] + o.send_off * EB, static_cast<size_t>(o.count * EB), b[0] == 0.0 && b[1] == 0.0, s);
                if (st == COSMA_B200_OK) st = peer_wait_arrived(peer, *link, s);
                if (st != COSMA_B200_OK) return st;
            }
            if (record_end) CUDA_TRY(cudaEventRecord(plan.micro_ev[2 * i + 1], s));
            last_comm = static_cast<int>(i);
            continue;
        }
        switch (o.kind) {
            case cosma::MicroKind::GEMM: {
                double b[2];
                beta_of(o.beta, beta, E, b);
                ScopedReservedSms guard(o.narrow ? plan.reserved : 0);
                int path = 0;
                st = launch_gemm_nn(plan.dtype, s, o.m, o.n, o.k, alpha, arenas[0] + o.a_off * EB, o.lda, arenas[1] + o.b_off * EB, o.ldb, b,
                                    arenas[2] + o.c_off * EB, o.ldc, &path);
                if (path) ++plan.last_launches;
                break;
            }
            case cosma::MicroKind::ALLGATHER:
                st = run_allgather(plan, ops[o.op], arenas[ops[o.op].matrix], s);
                break;
            case cosma::MicroKind::SERIAL: {
                const auto& op = ops[o.op];
                if (op.kind == cosma::OpKind::ALLGATHER) st = run_allgather(plan, op, arenas[op.matrix], s);
                else if (op.kind == cosma::OpKind::REDUCE) st = run_reduce(plan, op, arenas[op.matrix], beta, s);
                else st = COSMA_B200_INTERNAL_ERROR;  // a program holds exactly one GEMM, lowered into panels
                break;
            }
            case cosma::MicroKind::EXCHANGE: {
                ncclComm_t ring = plan.ring_comms[o.ring_index];
                NCCL_TRY(N->GroupStart());
                NCCL_TRY(N->Send(arenas[2] + o.send_off * EB, static_cast<size_t>(o.count) * E, ndt, o.peer, ring, s));
                double b[2];
                beta_of(o.beta, beta, E, b);
                const int64_t recv_off = (b[0] == 0.0 && b[1] == 0.0) ? o.recv_off_zero : o.recv_off;
                NCCL_TRY(N->Recv(arenas[2] + recv_off * EB, static_cast<size_t>(o.count) * E, ndt, o.peer, ring, s));
                NCCL_TRY(N->GroupEnd());
                break;
            }
            case cosma::MicroKind::ACCUMULATE: {
                double b[2];
                beta_of(o.beta, beta, E, b);
                if (o.beta_term && b[0] == 0.0 && b[1] == 0.0) break;  // the exchange landed in C itself
                st = launch_axpby(plan, o.count, b[0], b[1], arenas[2] + o.dst_off * EB, arenas[2] + o.add_off * EB, s);
                break;
            }
        }
        if (st != COSMA_B200_OK) return st;
        CUDA_TRY(cudaEventRecord(plan.micro_ev[2 * i + 1], s));
        if (o.stream) last_comm = static_cast<int>(i);
    }
    // the caller's stream is complete only when the communication stream is
    if (last_comm >= 0) CUDA_TRY(cudaStreamWaitEvent(stream, plan.micro_ev[2 * last_comm + 1], 0));
    return COSMA_B200_OK;
}

}  // namespace

int plan_run(Plan& plan, const double* alpha, const double* beta, void* A_, void* B_, void* C_, cudaStream_t stream,
             const HostOperands* host, unsigned skip_allgather_mask) {
    const int E = plan.elem_reals;
    const int64_t EB = plan.elem_bytes();
    char* A = static_cast<char*>(A_);
    char* B = static_cast<char*>(B_);
    char* C = static_cast<char*>(C_);
    char* arenas[3] = {A, B, C};
    plan.last_launches = 0;
    plan.last_run_overlapped = false;
    if (plan.overlap.enabled && !host && skip_allgather_mask == 0) {
        // Transports of the overlapped program: copy engines when these arenas are bound (the default), else NCCL kernels beside narrow
        // GEMMs -- which needs ring communicators limited to the SMs those GEMMs leave free (rings_capped: COSMA_B200_PEER_COPY=OFF
        // at plan creation). With neither, the serial schedule is the better one. The arenas are bound collectively and the verdict is
        // shared, so every rank of a ring decides alike.
        bool bound = false;
        for (const auto& t : plan.peers) bound |= t->ready && A == t->bound[0] && B == t->bound[1] && C == t->bound[2];
        if (bound || plan.rings_capped) {
            plan.last_run_overlapped = true;
            return plan_run_overlapped(plan, alpha, beta, arenas, stream);
        }
    }
    // with timing on, every op (GEMM, allgather, reduce) is bracketed by a pair of events: ev[2*i], ev[2*i+1] for op i
    if (plan.time_gemms) {
        while (plan.ev.size() < 2 * plan.schedule.ops().size()) {
            cudaEvent_t e;
            CUDA_TRY(cudaEventCreate(&e));
            plan.ev.push_back(e);
        }
    }
    size_t oi = 0;
    for (const auto& op : plan.schedule.ops()) {
        int st = COSMA_B200_OK;
        if (plan.time_gemms) CUDA_TRY(cudaEventRecord(plan.ev[2 * oi], stream));
        switch (op.kind) {
            case cosma::OpKind::GEMM: {
                double b[2] = {0.0, 0.0};
                if (op.beta == cosma::BetaMode::ONE) b[0] = 1.0;
                else if (op.beta == cosma::BetaMode::USER) { b[0] = beta[0]; b[1] = E == 2 ? beta[1] : 0.0; }
                const int64_t lda = std::max(op.m, 1), ldb = std::max(op.k, 1), ldc = std::max(op.m, 1);
                StreamGemmArgs g;
                g.dtype = plan.dtype;
                g.m = op.m; g.n = op.n; g.k = op.k;
                g.alpha = alpha; g.beta = b;
                g.dA = A + op.a_off * EB; g.dlda = lda;
                g.dB = B + op.b_off * EB; g.dldb = ldb;
                g.dC = C + op.c_off * EB; g.dldc = ldc;
                if (host) {  // operands of this (only) GEMM still in host memory: PCIe pipelined under the kernel
                    g.hA = host->A; g.lda = lda;
                    g.hB = host->B; g.ldb = ldb;
                    g.hC_in = host->C_in; g.hC_out = host->C_out; g.ldc = ldc;
                }
                int launches = 0;
                st = stream_gemm(stream, g, &launches);
                plan.last_launches += launches;
                break;
            }
            case cosma::OpKind::ALLGATHER:
                if (!((skip_allgather_mask >> op.matrix) & 1u)) st = run_allgather(plan, op, arenas[op.matrix], stream);
                break;
            case cosma::OpKind::REDUCE:
                st = run_reduce(plan, op, arenas[op.matrix], beta, stream);
                break;
        }
        if (st != COSMA_B200_OK) return st;
        if (plan.time_gemms) CUDA_TRY(cudaEventRecord(plan.ev[2 * oi + 1], stream));
        ++oi;
    }
    return COSMA_B200_OK;
}

bool host_panel_pieces(const cosma::Schedule& schedule, int rank, int c, int j, std::vector<PanelPiece>& b_pieces,
                       std::vector<PanelPiece>& c_pieces) {
    b_pieces.clear();
    c_pieces.clear();
    const int P = static_cast<int>(schedule.strategy().P);
    if (c < 1 || j < 0 || j >= c || rank < 0 || rank >= P) return false;
    // elementary column ranges: consecutive cuts of the union of all B and C block boundaries
    std::vector<int> cuts;
    for (int x = 1; x <= 2; ++x)
        for (int q = 0; q < P; ++q) {
            const auto& blocks = schedule.mapper(x).initial_layout(q);
            if (blocks.size() != 1) return false;
            cuts.push_back(blocks[0].cols.first());
            cuts.push_back(blocks[0].cols.last() + 1);
        }
    std::sort(cuts.begin(), cuts.end());
    cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
    for (size_t i = 0; i + 1 < cuts.size(); ++i)
        if ((cuts[i + 1] - cuts[i]) % c != 0) return false;
    for (int x = 1; x <= 2; ++x) {
        const auto& mine = schedule.mapper(x).initial_layout(rank)[0];
        const std::int64_t first = mine.cols.first(), last = mine.cols.last(), rows = mine.rows.length();
        std::vector<PanelPiece>& out = x == 1 ? b_pieces : c_pieces;
        std::int64_t pos = 0;
        for (size_t i = 0; i + 1 < cuts.size(); ++i) {
            if (cuts[i] < first || cuts[i + 1] > last + 1) continue;  // block boundaries are cuts: an elementary range is inside or outside
            const std::int64_t w = (cuts[i + 1] - cuts[i]) / c;
            out.push_back(PanelPiece{(cuts[i] - first + j * w) * rows, w * rows, pos});
            pos += w * rows;
        }
    }
    return true;
}

// COSMA_B200_HOST_PANELS=c (default 4; seen green on 2 and 8 GPUs, profiles/r2b_pytest_gpu_n2.txt, r2d_bench_n8_panels4.json): the host-pointer multiply of a
// schedule with collectives as c column panels. A goes up (and is gathered) once; panel j's pieces of B (and of C when beta != 0)
// travel up on a copy stream while panel j - 1 multiplies, panel j - 1's C travels down on another. Two B / C arena sets alternate.
// Every decision below depends on global information only (strategy, shapes, c), so all ranks take the same path.
int host_panels_requested() {
    // default 4: measured at 8 GPUs on 32768^3 (profiles/r2d_bench_n8_*.json): 206 TFLOP/s end to end against 165 with up-front copies
    // (device-resident: 290); COSMA_B200_HOST_PANELS=0 | 1 switches the panel pipeline off, c > 1 sets the panel count
    static const int c = [] {
        const char* v = std::getenv("COSMA_B200_HOST_PANELS");
        const int x = v && *v ? std::atoi(v) : 4;
        return x > 1 ? x : 0;
    }();
    return c;
}

#define PANEL_CUDA(call)                                                                  \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            set_last_error(std::string("host panels: ") + #call + ": " + cudaGetErrorString(e_)); \
            return COSMA_B200_CUDA_ERROR;                                                 \
        }                                                                                 \
    } while (0)

// *handled = false: the schedule cannot be cut into c panels (same verdict on every rank); the caller takes the plain path.
int multiply_host_panels(Plan* p, int c, const double* alpha, const double* beta, const void* A, const void* B, void* C, cudaStream_t st,
                         bool* handled) {
    *handled = false;
    if (p->panel_count < 0) return COSMA_B200_OK;  // found unsuitable before
    if (!p->panel_plan) {
        p->panel_count = -1;  // until proven suitable
        const cosma::Strategy& full = p->schedule.strategy();
        const int P = static_cast<int>(full.P);
        int n_gemm = 0;
        bool gathered[3] = {false, false, false};
        for (const auto& op : p->schedule.ops()) {
            if (op.kind == cosma::OpKind::GEMM) ++n_gemm;
            else gathered[op.matrix] = true;
        }
        // one GEMM, and an operand that is gathered first (otherwise the plain path already streams A and B under the kernel)
        if (P < 2 || n_gemm != 1 || !(gathered[0] || gathered[1]) || full.n % c != 0 || p->ring_comms.empty()) return COSMA_B200_OK;
        std::vector<PanelPiece> probe_b, probe_c;
        for (int q = 0; q < P; ++q)  // the verdict must not depend on the rank
            if (!host_panel_pieces(p->schedule, q, c, 0, probe_b, probe_c)) return COSMA_B200_OK;
        const std::string steps = full.to_string();
        std::unique_ptr<Plan> sp(new Plan);
        try {
            const cosma::Strategy::quiet_errors hush;  // an unsuitable panel shape is an answer, not news
            const cosma::Strategy sub = cosma::parse_strategy(full.m, full.n / c, full.k, static_cast<size_t>(P), steps);
            if (sub.to_string() != steps) return COSMA_B200_OK;
            sp->schedule = cosma::Schedule(sub, p->schedule.rank());
        } catch (const std::exception&) {
            return COSMA_B200_OK;
        }
        sp->dtype = p->dtype; sp->elem_reals = p->elem_reals; sp->real_bytes = p->real_bytes;
        const auto& ra = p->schedule.rings();
        const auto& rb = sp->schedule.rings();
        bool same = ra.size() == rb.size();
        for (size_t i = 0; same && i < ra.size(); ++i) same = ra[i].step == rb[i].step && ra[i].color == rb[i].color && ra[i].my_pos == rb[i].my_pos && ra[i].ranks == rb[i].ranks;
        if (!same || sp->schedule.initial_elements(0) != p->schedule.initial_elements(0) || sp->schedule.arena_elements(0) > p->schedule.arena_elements(0) ||
            sp->schedule.initial_elements(1) * c != p->schedule.initial_elements(1) || sp->schedule.initial_elements(2) * c != p->schedule.initial_elements(2))
            return COSMA_B200_OK;
        sp->ring_comms = p->ring_comms;  // the same rings: borrowed, destroyed with the parent
        sp->borrowed_comms = true;
        p->panel_plan = sp.release();
        p->panel_count = c;
    }
    if (p->panel_count != c) return COSMA_B200_INVALID_ARG;  // one panel count per plan
    std::vector<PanelPiece> pieces, cpieces;
    Plan* sp = p->panel_plan;
    const size_t es = static_cast<size_t>(p->elem_bytes());
    // arenas: A shared by all panels (large enough for either plan), two sets of the panel plan's B and C arenas
    const size_t a_bytes = std::max<size_t>(p->schedule.arena_elements(0), 1) * es;
    if (!p->owned[0]) PANEL_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->owned[0]), a_bytes));
    for (int s = 0; s < 2; ++s)
        for (int x = 0; x < 2; ++x)
            if (!p->panel_arena[s][x])
                PANEL_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->panel_arena[s][x]), std::max<size_t>(sp->schedule.arena_elements(1 + x), 1) * es));
    if (!p->panel_in) PANEL_CUDA(cudaStreamCreateWithFlags(&p->panel_in, cudaStreamNonBlocking));
    if (!p->panel_out) PANEL_CUDA(cudaStreamCreateWithFlags(&p->panel_out, cudaStreamNonBlocking));
    while (p->panel_ev.size() < static_cast<size_t>(3 + 3 * c)) {
        cudaEvent_t e;
        PANEL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        p->panel_ev.push_back(e);
    }
    cudaEvent_t ev_start = p->panel_ev[0], ev_a = p->panel_ev[1], ev_end = p->panel_ev[2];
    cudaEvent_t* ev_in = p->panel_ev.data() + 3;  // [c] panel j's operands are on the device
    cudaEvent_t* ev_done = ev_in + c;             // [c] panel j is computed
    cudaEvent_t* ev_out = ev_done + c;            // [c] panel j's C is back on the host
    const bool beta_zero = beta[0] == 0.0 && (p->elem_reals == 1 || beta[1] == 0.0);
    const char* hB = static_cast<const char*>(B);
    char* hC = static_cast<char*>(C);
    cudaStream_t cin = p->panel_in, cout = p->panel_out;
    // the copy streams start after whatever the caller queued on `st` (earlier users of the arenas)
    PANEL_CUDA(cudaEventRecord(ev_start, st));
    PANEL_CUDA(cudaStreamWaitEvent(cin, ev_start, 0));
    PANEL_CUDA(cudaStreamWaitEvent(cout, ev_start, 0));
    const size_t a_init = static_cast<size_t>(p->schedule.initial_elements(0)) * es;
    if (a_init) PANEL_CUDA(cudaMemcpyAsync(p->owned[0], A, a_init, cudaMemcpyHostToDevice, cin));
    PANEL_CUDA(cudaEventRecord(ev_a, cin));
    p->last_launches = 0;
    for (int j = 0; j < c; ++j) {
        const int s = j & 1;
        char* dB = p->panel_arena[s][0];
        char* dC = p->panel_arena[s][1];
        if (!host_panel_pieces(p->schedule, p->schedule.rank(), c, j, pieces, cpieces)) return COSMA_B200_INTERNAL_ERROR;
        if (j >= 2) {  // set s was last used by panel j - 2: its GEMM read B, its download read C
            PANEL_CUDA(cudaStreamWaitEvent(cin, ev_done[j - 2], 0));
            PANEL_CUDA(cudaStreamWaitEvent(cin, ev_out[j - 2], 0));
            PANEL_CUDA(cudaStreamWaitEvent(st, ev_out[j - 2], 0));
        }
        for (const auto& pc : pieces)
            PANEL_CUDA(cudaMemcpyAsync(dB + pc.dst_off * es, hB + pc.src_off * es, static_cast<size_t>(pc.len) * es, cudaMemcpyHostToDevice, cin));
        if (!beta_zero)
            for (const auto& pc : cpieces)
                PANEL_CUDA(cudaMemcpyAsync(dC + pc.dst_off * es, hC + pc.src_off * es, static_cast<size_t>(pc.len) * es, cudaMemcpyHostToDevice, cin));
        PANEL_CUDA(cudaEventRecord(ev_in[j], cin));
        if (j == 0) PANEL_CUDA(cudaStreamWaitEvent(st, ev_a, 0));
        PANEL_CUDA(cudaStreamWaitEvent(st, ev_in[j], 0));
        const int rc = plan_run(*sp, alpha, beta, p->owned[0], dB, dC, st, nullptr, j == 0 ? 0u : 1u /* A stays gathered */);
        if (rc != COSMA_B200_OK) return rc;
        p->last_launches += sp->last_launches;
        PANEL_CUDA(cudaEventRecord(ev_done[j], st));
        PANEL_CUDA(cudaStreamWaitEvent(cout, ev_done[j], 0));
        for (const auto& pc : cpieces)
            PANEL_CUDA(cudaMemcpyAsync(hC + pc.src_off * es, dC + pc.dst_off * es, static_cast<size_t>(pc.len) * es, cudaMemcpyDeviceToHost, cout));
        PANEL_CUDA(cudaEventRecord(ev_out[j], cout));
    }
    // the caller's stream completes only when the last panel is back on the host
    PANEL_CUDA(cudaEventRecord(ev_end, cout));
    PANEL_CUDA(cudaStreamWaitEvent(st, ev_end, 0));
    *handled = true;
    return COSMA_B200_OK;
}
#undef PANEL_CUDA

}  // namespace cosma_b200

using cosma_b200::Comm;
using cosma_b200::Plan;
using cosma_b200::nccl;
using cosma_b200::set_last_error;
using cosma_b200::guarded;

extern "C" {

int cosma_b200_nccl_unique_id(uint8_t* out128) {
    return guarded("cosma_b200_nccl_unique_id", [&]() -> int {
        const auto* N = nccl();
        if (!N) return COSMA_B200_NCCL_ERROR;
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
        ncclUniqueId id;
        if (N->GetUniqueId(&id) != ncclSuccess) return COSMA_B200_NCCL_ERROR;
        std::memcpy(out128, &id, 128);
        return COSMA_B200_OK;
    });
}

int cosma_b200_comm_create(int rank, int nranks, const uint8_t* id128, void** comm_out) {
    return guarded("cosma_b200_comm_create", [&]() -> int {
        if (!comm_out || nranks < 1 || rank < 0 || rank >= nranks) return COSMA_B200_INVALID_ARG;
        auto c = std::make_unique<Comm>();
        c->rank = rank;
        c->size = nranks;
        if (nranks == 1) {  // single-rank job: nothing to exchange, NCCL is not needed (id128 may be NULL)
            *comm_out = c.release();
            return COSMA_B200_OK;
        }
        const auto* N = nccl();
        if (!N || !id128) return COSMA_B200_NCCL_ERROR;
        ncclUniqueId id;
        std::memcpy(&id, id128, 128);
        ncclResult_t r = N->CommInitRank(&c->comm, nranks, id, rank);
        if (r != ncclSuccess) {
            set_last_error(std::string("ncclCommInitRank: ") + N->GetErrorString(r));
            return COSMA_B200_NCCL_ERROR;
        }
        *comm_out = c.release();
        return COSMA_B200_OK;
    });
}

int cosma_b200_comm_destroy(void* comm) {
    Comm* c = static_cast<Comm*>(comm);
    if (!c) return COSMA_B200_OK;
    if (c->comm && nccl()) nccl()->CommDestroy(c->comm);
    delete c;
    return COSMA_B200_OK;
}

static int plan_create_impl(void* comm, int rank, int nranks, int m, int n, int k, int P, const char* steps, char dtype, void** plan_out);

int cosma_b200_plan_create(void* comm, int rank, int nranks, int m, int n, int k, const char* steps, char dtype,
                           void** plan_out) {
    return plan_create_impl(comm, rank, nranks, m, n, k, /*P=*/0, steps, dtype, plan_out);
}
int cosma_b200_plan_create_for_strategy(void* comm, int rank, int nranks, int m, int n, int k, int P, const char* steps, char dtype,
                                        void** plan_out) {
    if (P < 1) {
        set_last_error("plan_create_for_strategy: P must be >= 1");
        return COSMA_B200_INVALID_ARG;
    }
    return plan_create_impl(comm, rank, nranks, m, n, k, P, steps, dtype, plan_out);
}

// P == 0: Strategy(m, n, k, nranks) completed automatically when `steps` is empty; P >= 1: exactly the caller's strategy
static int plan_create_impl(void* comm, int rank, int nranks, int m, int n, int k, int P, const char* steps, char dtype, void** plan_out) {
    try {
        if (dtype != 'd' && dtype != 'z' && dtype != 's' && dtype != 'c') {
            set_last_error("plan dtype must be one of s, d, c, z");
            return COSMA_B200_INVALID_ARG;
        }
        Comm* c = static_cast<Comm*>(comm);
        if (c) { rank = c->rank; nranks = c->size; }
        if (P > nranks) {
            set_last_error("the strategy uses more ranks than the communicator has");
            return COSMA_B200_INVALID_ARG;
        }
        cosma::Strategy strategy;
        if (P == 0) {
            // automatic / prefixed strategy: COSMA_CPU_MAX_MEMORY and COSMA_B200_DEVICE_MEMORY_MB apply (auto_strategy.hpp)
            strategy = cosma::automatic_strategy(m, n, k, nranks, steps ? steps : "", (dtype == 'd' || dtype == 'z' ? 8 : 4) * (dtype == 'z' || dtype == 'c' ? 2 : 1));
        } else {
            const std::string st = steps ? steps : "";
            if (st.find_first_not_of(" ,") == std::string::npos) {
                std::vector<int> divs;
                std::string dims, types;
                strategy = cosma::Strategy(m, n, k, static_cast<size_t>(P), divs, dims, types);  // valid only for P == 1
            } else {
                strategy = cosma::parse_strategy(m, n, k, static_cast<size_t>(P), st);
            }
        }
        // a dimension cut into more parts than it has elements: the reference accepts it and then computes garbage (Interval::subinterval
        // hands out the whole interval when the divisor exceeds the length, interval.cpp:84-98); refuse instead of multiplying wrongly
        {
            long long parts[3] = {1, 1, 1};
            for (size_t s = 0; s < strategy.n_steps(); ++s) parts[strategy.split_m(s) ? 0 : (strategy.split_n(s) ? 1 : 2)] *= strategy.divisor(s);
            const long long dims[3] = {m, n, k};
            for (int d = 0; d < 3; ++d)
                if (parts[d] > dims[d] && dims[d] > 0) {
                    set_last_error(std::string("the strategy divides dimension ") + "mnk"[d] + " = " + std::to_string(dims[d]) + " into " +
                                   std::to_string(parts[d]) + " parts: not a meaningful decomposition (the reference computes a wrong product here)");
                    return COSMA_B200_INVALID_ARG;
                }
        }
        auto plan = std::make_unique<Plan>();
        plan->schedule = cosma::Schedule(strategy, rank);
        plan->dtype = dtype;
        plan->elem_reals = (dtype == 'z' || dtype == 'c') ? 2 : 1;
        plan->real_bytes = (dtype == 'd' || dtype == 'z') ? 8 : 4;
        bool use_ring_config = false;
        // communication / computation overlap: lower the tail of the op list if EVERY active rank's schedule allows it (ring mates must
        // speak the same protocol; every rank evaluates the same ranks' schedules, so all reach the same verdict)
        {
            int sms = 0, dev = 0;
            if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 0;
            (void)cudaGetLastError();
            const cosma::OverlapTuning tuning = cosma::overlap_tuning_from_env(dtype, sms);
            const int Pn = static_cast<int>(strategy.P);
            bool all = tuning.enabled && Pn > 1 && Pn <= 64;
            std::string why;
            for (int r = 0; all && r < Pn; ++r) {
                cosma::OverlapProgram pr = r == rank ? cosma::plan_overlap(plan->schedule, tuning) : cosma::plan_overlap(cosma::Schedule(strategy, r), tuning);
                if (!pr.enabled) { all = false; why = "rank " + std::to_string(r) + ": " + pr.why; }
                if (r == rank) plan->overlap = std::move(pr);
            }
            if (all) {
                plan->reserved = tuning.reserved_sms;  // idle ranks too: they take part in the communicator splits with the same configuration
            } else {
                plan->overlap = cosma::OverlapProgram();
                plan->overlap.why = !why.empty() ? why : (!tuning.enabled ? "switched off (COSMA_OVERLAP_COMM_AND_COMP)" : "not a multi-rank schedule of at most 64 ranks");
            }
            // Only the NCCL transport of the overlapped program needs capped rings. With the copy-engine transport (the default) the rings
            // carry the SERIAL collectives -- the host-pointer paths (column panels, streamed operands) and rings larger than two --
            // and those want every channel NCCL can use: 8 CTAs move ~89 GB/s (profiles/r2b_bench_n2_*.json) against 460-490 uncapped.
            use_ring_config = all && !cosma_b200::peer_copy_enabled();
            plan->rings_capped = use_ring_config;
            plan->overlap_job = all;
        }
        plan->parent = c;
        if (c && nranks > 1) {
            const auto* N = nccl();
            // overlapped plans: the ring communicators' kernels are limited to the SMs the narrow GEMMs leave free
            ncclConfig_t ring_config = NCCL_CONFIG_INITIALIZER;
            ring_config.minCTAs = 1;
            ring_config.maxCTAs = std::max(1, plan->reserved);
            ncclConfig_t* config = use_ring_config ? &ring_config : nullptr;
            // one communicator split per parallel step, called by EVERY rank of the parent communicator in step
            // order (idle ranks and ranks outside a ring pass NCCL_SPLIT_NOCOLOR)
            std::vector<int> par_steps;
            for (size_t s = 0; s < strategy.n_steps(); ++s)
                if (strategy.parallel_step(s)) par_steps.push_back(static_cast<int>(s));
            const auto& rings = plan->schedule.rings();
            plan->ring_comms.assign(rings.size(), nullptr);
            for (int s : par_steps) {
                int color = NCCL_SPLIT_NOCOLOR, key = 0, idx = -1;
                for (size_t i = 0; i < rings.size(); ++i)
                    if (rings[i].step == s) { color = rings[i].color; key = rings[i].my_pos; idx = static_cast<int>(i); }
                ncclComm_t sub = nullptr;
                ncclResult_t r = N->CommSplit(c->comm, color, key, &sub, config);
                if (r != ncclSuccess) {
                    set_last_error(std::string("ncclCommSplit: ") + N->GetErrorString(r));
                    cosma_b200_plan_destroy(plan.release());  // also destroys the ring communicators split so far
                    return COSMA_B200_NCCL_ERROR;
                }
                if (idx >= 0) plan->ring_comms[idx] = sub;
            }
        }
        *plan_out = plan.release();
        return COSMA_B200_OK;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
}

int cosma_b200_plan_destroy(void* plan) {
    Plan* p = static_cast<Plan*>(plan);
    if (!p) return COSMA_B200_OK;
    if (!p->borrowed_comms)
        for (auto c : p->ring_comms)
            if (c && nccl()) nccl()->CommDestroy(c);
    for (auto e : p->ev) cudaEventDestroy(e);
    for (auto& t : p->peers) cosma_b200::peer_transport_release(*t);
    p->peers.clear();
    for (auto e : p->micro_ev) cudaEventDestroy(e);
    if (p->comm_stream) cudaStreamDestroy(p->comm_stream);
    for (auto& a : p->owned)
        if (a) cudaFree(a);
    for (auto e : p->panel_ev) cudaEventDestroy(e);
    for (auto& set : p->panel_arena)
        for (auto& a : set)
            if (a) cudaFree(a);
    if (p->panel_in) cudaStreamDestroy(p->panel_in);
    if (p->panel_out) cudaStreamDestroy(p->panel_out);
    if (p->panel_plan) cosma_b200_plan_destroy(p->panel_plan);
    delete p;
    return COSMA_B200_OK;
}

static bool valid_matrix(int matrix) { return matrix >= 0 && matrix <= 2; }

int64_t cosma_b200_plan_arena_elements(void* plan, int matrix) {
    if (!plan || !valid_matrix(matrix)) return -1;
    return static_cast<Plan*>(plan)->schedule.arena_elements(matrix);
}
int64_t cosma_b200_plan_initial_elements(void* plan, int matrix) {
    if (!plan || !valid_matrix(matrix)) return -1;
    return static_cast<Plan*>(plan)->schedule.initial_elements(matrix);
}
int cosma_b200_plan_strategy(void* plan, char* out, int out_len, int* P_used) {
    if (!plan || !out) return COSMA_B200_INVALID_ARG;
    return guarded("cosma_b200_plan_strategy", [&] {
        const auto& st = static_cast<Plan*>(plan)->schedule.strategy();
        const std::string s = st.to_string();
        if (static_cast<int>(s.size()) + 1 > out_len) return static_cast<int>(COSMA_B200_INVALID_ARG);
        std::strcpy(out, s.c_str());
        if (P_used) *P_used = static_cast<int>(st.P);
        return static_cast<int>(COSMA_B200_OK);
    });
}
double cosma_b200_plan_gemm_flops(void* plan) {
    Plan* p = static_cast<Plan*>(plan);
    if (!p) return 0.0;
    return p->schedule.total_gemm_flops() * (p->elem_reals == 2 ? 4.0 : 1.0);
}
int cosma_b200_plan_export(void* plan, int64_t* buf, int64_t cap, int64_t* len) {
    if (!plan || !len) return COSMA_B200_INVALID_ARG;
    return guarded("cosma_b200_plan_export", [&] {
        const auto v = static_cast<Plan*>(plan)->schedule.serialize();
        *len = static_cast<int64_t>(v.size());
        if (buf && cap >= *len) std::memcpy(buf, v.data(), v.size() * sizeof(int64_t));
        return static_cast<int>(COSMA_B200_OK);
    });
}
int cosma_b200_plan_local_blocks(void* plan, int matrix, int rank, int* out, int cap, int* n_blocks) {
    if (!plan || !n_blocks || !valid_matrix(matrix)) return COSMA_B200_INVALID_ARG;
    return guarded("cosma_b200_plan_local_blocks", [&] {
        const cosma::Schedule& sch = static_cast<Plan*>(plan)->schedule;
        if (rank < 0 || rank >= static_cast<int>(sch.strategy().P)) {  // ranks the strategy leaves idle own nothing
            *n_blocks = 0;
            return static_cast<int>(COSMA_B200_OK);
        }
        const auto& blocks = sch.mapper(matrix).initial_layout(rank);
        *n_blocks = static_cast<int>(blocks.size());
        if (out && cap >= 4 * *n_blocks)
            for (int i = 0; i < *n_blocks; ++i) {
                out[4 * i] = blocks[i].rows.first(); out[4 * i + 1] = blocks[i].rows.last();
                out[4 * i + 2] = blocks[i].cols.first(); out[4 * i + 3] = blocks[i].cols.last();
            }
        return static_cast<int>(COSMA_B200_OK);
    });
}

/* Planning only: column panel j of c of this plan's rank (exec_internal.h host_panel_pieces). b_pieces / c_pieces: (src_off, len, dst_off)
 * triples of local B / local C. *eligible = 0 when the layout cannot be cut this way. */
int cosma_b200_plan_host_panel(void* plan, int c, int j, int64_t* b_pieces, int b_cap, int* n_b, int64_t* c_pieces, int c_cap, int* n_c,
                               int* eligible) {
    if (!plan || !n_b || !n_c || !eligible) return COSMA_B200_INVALID_ARG;
    return guarded("cosma_b200_plan_host_panel", [&]() -> int {
        Plan* p = static_cast<Plan*>(plan);
        std::vector<cosma_b200::PanelPiece> vb, vc;
        *n_b = 0; *n_c = 0;
        *eligible = !p->schedule.idle() && cosma_b200::host_panel_pieces(p->schedule, p->schedule.rank(), c, j, vb, vc) ? 1 : 0;
        if (!*eligible) return COSMA_B200_OK;
        *n_b = static_cast<int>(vb.size());
        *n_c = static_cast<int>(vc.size());
        if (b_pieces && b_cap >= 3 * *n_b)
            for (size_t i = 0; i < vb.size(); ++i) { b_pieces[3 * i] = vb[i].src_off; b_pieces[3 * i + 1] = vb[i].len; b_pieces[3 * i + 2] = vb[i].dst_off; }
        if (c_pieces && c_cap >= 3 * *n_c)
            for (size_t i = 0; i < vc.size(); ++i) { c_pieces[3 * i] = vc[i].src_off; c_pieces[3 * i + 1] = vc[i].len; c_pieces[3 * i + 2] = vc[i].dst_off; }
        return COSMA_B200_OK;
    });
}

int cosma_b200_multiply(void* plan, const double* alpha, const double* beta, void* A, void* B, void* C, void* stream) {
    return guarded("cosma_b200_multiply", [&]() -> int {
        Plan* p = static_cast<Plan*>(plan);
        if (!p || !alpha || !beta) return COSMA_B200_INVALID_ARG;
        if (p->schedule.idle()) return COSMA_B200_OK;
        bool needs_comm = false;
        for (const auto& op : p->schedule.ops()) needs_comm |= op.kind != cosma::OpKind::GEMM;
        if (needs_comm && p->ring_comms.empty()) {
            set_last_error("plan was created without a communicator (plan-only); cannot execute collectives");
            return COSMA_B200_INVALID_ARG;
        }
        return cosma_b200::plan_run(*p, alpha, beta, A, B, C, static_cast<cudaStream_t>(stream), nullptr);
    });
}

/* Host-pointer variant: local A, B (and C when beta != 0) are uploaded from (pinned) host memory into arenas owned by
 * the plan, the schedule runs from HBM, and local C is downloaded; everything is ordered on `stream`. This is the
 * calling convention of the reference, whose CosmaMatrix buffers live in host memory even in GPU/NCCL builds
 * (src/cosma/local_multiply.cpp:341-363, gpu/nccl_utils.cpp:98,124-135). */
int cosma_b200_multiply_host(void* plan, const double* alpha, const double* beta, const void* A, const void* B, void* C,
                             void* stream) {
    return guarded("cosma_b200_multiply_host", [&]() -> int {
        Plan* p = static_cast<Plan*>(plan);
        if (!p || !alpha || !beta) return COSMA_B200_INVALID_ARG;
        if (p->schedule.idle()) {
            // idle ranks take part in the one collective of this entry point: the verdict of the arena binding (first call)
            if (p->overlap_job && !p->owned_bound && cosma_b200::host_panels_requested() == 0) {
                p->owned_bound = true;
                return cosma_b200_plan_bind_arenas(p, nullptr, nullptr, nullptr, nullptr);
            }
            return COSMA_B200_OK;
        }
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (const int panels = cosma_b200::host_panels_requested()) {
            bool handled = false;
            const int rc = cosma_b200::multiply_host_panels(p, panels, alpha, beta, A, B, C, st, &handled);
            if (rc != COSMA_B200_OK || handled) return rc;
        }
        const size_t es = static_cast<size_t>(p->elem_bytes());
        for (int x = 0; x < 3; ++x)
            if (!p->owned[x]) {
                const size_t bytes = std::max<size_t>(p->schedule.arena_elements(x), 1) * es;
                if (cudaMalloc(reinterpret_cast<void**>(&p->owned[x]), bytes) != cudaSuccess) {
                    set_last_error("cudaMalloc of a plan arena failed");
                    return COSMA_B200_OUT_OF_MEMORY;
                }
            }
        // the plan's own arenas: copy-engine transport for the overlapped transfers (collective, idle ranks included -- hence a condition
        // that every rank evaluates alike; with the panel pipeline, the default, the overlapped path is not used from here)
        if (p->overlap_job && !p->owned_bound && cosma_b200::host_panels_requested() == 0) {
            p->owned_bound = true;
            const int rc = cosma_b200_plan_bind_arenas(p, p->owned[0], p->owned[1], p->owned[2], nullptr);
            if (rc != COSMA_B200_OK) return rc;
        }
        const bool beta_zero = beta[0] == 0.0 && (p->elem_reals == 1 || beta[1] == 0.0);
        // A schedule with ONE base-case GEMM that reads a local matrix as the caller holds it (no allgather of that matrix
        // before it) / leaves local C as the caller wants it (no reduce after it) streams that matrix over PCIe under the
        // kernel instead of copying it up front (host_gemm.cu): P = 1, and the un-gathered operands of P > 1 strategies
        // (pk2 at P = 2: A and B; pn2,pk2 at P = 4: B; pk8 of the large-K config: A and B).
        const cosma::ScheduleOp* gemm_op = nullptr;
        int n_gemm = 0;
        bool gathered[3] = {false, false, false};
        for (const auto& op : p->schedule.ops()) {
            if (op.kind == cosma::OpKind::GEMM) { gemm_op = &op; ++n_gemm; }
            else gathered[op.matrix] = true;
        }
        cosma_b200::HostOperands hs;
        bool streamed[3] = {false, false, false};
        if (n_gemm == 1) {
            const auto& g = *gemm_op;
            const int64_t need[3] = {int64_t(g.m) * g.k, int64_t(g.k) * g.n, int64_t(g.m) * g.n};
            const int64_t off[3] = {g.a_off, g.b_off, g.c_off};
            for (int x = 0; x < 3; ++x)
                streamed[x] = !gathered[x] && off[x] == 0 && need[x] == p->schedule.initial_elements(x) && need[x] > 0;
            if (streamed[0]) hs.A = A;
            if (streamed[1]) hs.B = B;
            if (streamed[2]) { hs.C_in = C; hs.C_out = C; }
        }
        const void* host_in[3] = {A, B, C};
        for (int x = 0; x < 3; ++x) {
            if (streamed[x] || (x == 2 && beta_zero)) continue;
            const size_t bytes = p->schedule.initial_elements(x) * es;
            if (bytes && cudaMemcpyAsync(p->owned[x], host_in[x], bytes, cudaMemcpyHostToDevice, st) != cudaSuccess)
                return COSMA_B200_CUDA_ERROR;
        }
        bool needs_comm = false;
        for (const auto& op : p->schedule.ops()) needs_comm |= op.kind != cosma::OpKind::GEMM;
        if (needs_comm && p->ring_comms.empty()) {
            set_last_error("plan was created without a communicator (plan-only); cannot execute collectives");
            return COSMA_B200_INVALID_ARG;
        }
        const bool any = streamed[0] || streamed[1] || streamed[2];
        int rc = cosma_b200::plan_run(*p, alpha, beta, p->owned[0], p->owned[1], p->owned[2], st, any ? &hs : nullptr);
        if (rc != COSMA_B200_OK) return rc;
        const size_t cbytes = p->schedule.initial_elements(2) * es;
        if (!streamed[2] && cbytes && cudaMemcpyAsync(C, p->owned[2], cbytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) return COSMA_B200_CUDA_ERROR;
        return COSMA_B200_OK;
    });
}

/* Binds the device arenas the plan will be run on (collective over the plan's communicator: every rank calls it, idle ranks too). An
 * overlapped plan (cosma_b200_plan_overlap_export) then moves its ring-of-two transfers with COPY ENGINES straight into the ring
 * mates' arenas (CUDA IPC) instead of NCCL kernels, and re-plans its panels for a device that no longer shares SMs with communication.
 * *active = 1 when that transport is in place (on every rank alike); 0: plan not overlapped, COSMA_B200_PEER_COPY=OFF (overlap over NCCL),
 * or some rank could not map its mate's memory (serial schedule). cosma_b200_multiply must then be called with exactly these arenas. */
int cosma_b200_plan_bind_arenas(void* plan, void* A, void* B, void* C, int* active) {
    return guarded("cosma_b200_plan_bind_arenas", [&]() -> int {
        Plan* p = static_cast<Plan*>(plan);
        if (active) *active = 0;
        if (!p) return COSMA_B200_INVALID_ARG;
        if (!p->overlap_job || !p->parent || !cosma_b200::peer_copy_enabled()) return COSMA_B200_OK;
        if (!p->schedule.idle() && (!A || !B || !C)) return COSMA_B200_INVALID_ARG;
        COSMA_B200_CUDA_TRY(cudaDeviceSynchronize());  // an earlier multiply on the previous transport may still be running
        bool ok = false;
        // one transport per set of arenas; a set bound again is set up afresh, at most four sets are kept (oldest dropped first; the
        // same on every rank)
        std::unique_ptr<cosma_b200::PeerTransport> fresh(new cosma_b200::PeerTransport);
        for (size_t i = 0; i < p->peers.size(); ++i)
            if (p->peers[i]->bound[0] == A && p->peers[i]->bound[1] == B && p->peers[i]->bound[2] == C) {
                cosma_b200::peer_transport_release(*p->peers[i]);
                p->peers.erase(p->peers.begin() + i);
                break;
            }
        if (p->peers.size() >= 4) {
            cosma_b200::peer_transport_release(*p->peers.front());
            p->peers.erase(p->peers.begin());
        }
        const int st = cosma_b200::peer_transport_setup(*p, *fresh, p->parent, A, B, C, &ok);
        if (st != COSMA_B200_OK) return st;
        fresh->bound[0] = A; fresh->bound[1] = B; fresh->bound[2] = C;  // remembered also when the set-up was refused
        p->peers.push_back(std::move(fresh));
        cosma_b200::PeerTransport& bound_now = *p->peers.back();
        if (ok && p->overlap.enabled) {
            // the same lowering for a transport that costs no SM: every panel on the whole device, transfers at the NVLink copy rate
            int sms = 0, dev = 0;
            if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 0;
            cosma::OverlapTuning tuning = cosma::overlap_tuning_from_env(p->dtype, sms);
            tuning.zero_sm = true;
            tuning.force = true;  // the verdict "every rank lowers" has been taken at plan creation
            if (!std::getenv("COSMA_B200_OVERLAP_GBPS")) tuning.link_gbps = 550.0;  // peer copy over NVLink 5 (measured on this pool: ~770 one way)
            cosma::OverlapProgram prog = cosma::plan_overlap(p->schedule, tuning);
            if (!prog.enabled) {
                set_last_error("bind_arenas: the overlapped program could not be rebuilt: " + prog.why);
                return COSMA_B200_INTERNAL_ERROR;
            }
            // the transport's links name micro-ops by index: the communication ops keep their order, so re-index them
            std::vector<int> comm_ops;
            for (size_t i = 0; i < prog.ops.size(); ++i)
                if (prog.ops[i].stream == 1 && (prog.ops[i].kind == cosma::MicroKind::ALLGATHER || prog.ops[i].kind == cosma::MicroKind::EXCHANGE)) comm_ops.push_back(static_cast<int>(i));
            // (every transport of the plan: the earlier ones were indexed against the program they were set up with)
            for (auto& t : p->peers) {
                if (!t->ready) continue;
                if (comm_ops.size() != t->links.size()) return COSMA_B200_INTERNAL_ERROR;
                for (size_t l = 0; l < comm_ops.size(); ++l) t->links[l].micro = comm_ops[l];
            }
            (void)bound_now;
            prog.why += " (copy-engine peer transport)";
            p->overlap = std::move(prog);
            for (auto e : p->micro_ev) cudaEventDestroy(e);
            p->micro_ev.clear();
        }
        if (active) *active = ok ? 1 : 0;
        return COSMA_B200_OK;
    });
}

int cosma_b200_plan_last_launches(void* plan) { return static_cast<Plan*>(plan)->last_launches; }

int cosma_b200_plan_time_gemms(void* plan, int enable) {
    static_cast<Plan*>(plan)->time_gemms = enable != 0;
    return COSMA_B200_OK;
}
/* after a synchronised run with timing enabled: per-GEMM device milliseconds (overlapped plans: one entry per GEMM panel) */
int cosma_b200_plan_gemm_times(void* plan, float* out, int cap, int* n) {
    return guarded("cosma_b200_plan_gemm_times", [&]() -> int {
        Plan* p = static_cast<Plan*>(plan);
        if (!p || !n) return COSMA_B200_INVALID_ARG;
        if (p->last_run_overlapped) {
            const auto& prog = p->overlap.ops;
            int g = 0;
            for (const auto& o : prog) g += o.kind == cosma::MicroKind::GEMM;
            *n = g;
            if (!p->time_gemms || p->micro_ev.size() < 2 * prog.size()) return COSMA_B200_INVALID_ARG;
            g = 0;
            for (size_t i = 0; i < prog.size(); ++i) {
                if (prog[i].kind != cosma::MicroKind::GEMM) continue;
                if (g < cap && cudaEventElapsedTime(&out[g], p->micro_ev[2 * i], p->micro_ev[2 * i + 1]) != cudaSuccess) return COSMA_B200_CUDA_ERROR;
                ++g;
            }
            return COSMA_B200_OK;
        }
        const auto& ops = p->schedule.ops();
        int cnt = 0;
        for (const auto& op : ops) cnt += op.kind == cosma::OpKind::GEMM;
        *n = cnt;
        if (!p->time_gemms || p->ev.size() < 2 * ops.size()) return COSMA_B200_INVALID_ARG;
        int g = 0;
        for (size_t i = 0; i < ops.size(); ++i) {
            if (ops[i].kind != cosma::OpKind::GEMM) continue;
            if (g < cap && cudaEventElapsedTime(&out[g], p->ev[2 * i], p->ev[2 * i + 1]) != cudaSuccess) return COSMA_B200_CUDA_ERROR;
            ++g;
        }
        return COSMA_B200_OK;
    });
}

/* after a synchronised run with timing enabled: for every op of the schedule its kind (0 GEMM, 1 allgather, 2 reduce / exchange of the
 * partial C, 3 accumulate), device milliseconds on the stream it ran on, and for collectives the bytes this rank puts on / takes off the
 * wire: (d-1)/d of the gathered (allgather) or reduced (reduce-scatter) buffer, d = ring size -- the "bus bandwidth" convention of
 * SURVEY 8d. Overlapped plans report their micro-ops (GEMM panels, allgathers and exchange on the communication stream). */
int cosma_b200_plan_op_times(void* plan, int* kinds, float* ms, int64_t* wire_bytes, int cap, int* n) {
    return guarded("cosma_b200_plan_op_times", [&]() -> int {
        Plan* p = static_cast<Plan*>(plan);
        if (!p || !n) return COSMA_B200_INVALID_ARG;
        const auto& ops = p->schedule.ops();
        auto wire_of = [&](const cosma::ScheduleOp& op) {
            int64_t total = 0, mine = 0;
            if (op.kind != cosma::OpKind::GEMM) {
                for (size_t g = 0; g < op.piece.size(); ++g)
                    for (auto v : op.piece[g]) {
                        total += v;
                        if (static_cast<int>(g) == op.my_pos) mine += v;
                    }
            }
            return (total - mine) * p->elem_bytes();
        };
        auto kind_of = [](const cosma::ScheduleOp& op) { return op.kind == cosma::OpKind::GEMM ? 0 : (op.kind == cosma::OpKind::ALLGATHER ? 1 : 2); };
        if (p->last_run_overlapped) {
            const auto& prog = p->overlap.ops;
            *n = static_cast<int>(prog.size());
            if (!p->time_gemms || p->micro_ev.size() < 2 * prog.size()) return COSMA_B200_INVALID_ARG;
            for (size_t i = 0; i < prog.size() && static_cast<int>(i) < cap; ++i) {
                const auto& o = prog[i];
                int kind = 0;
                int64_t wire = 0;
                switch (o.kind) {
                    case cosma::MicroKind::GEMM: kind = 0; break;
                    case cosma::MicroKind::ALLGATHER:
                    case cosma::MicroKind::SERIAL: kind = kind_of(ops[o.op]); wire = wire_of(ops[o.op]); break;
                    case cosma::MicroKind::EXCHANGE: kind = 2; wire = o.count * p->elem_bytes(); break;
                    case cosma::MicroKind::ACCUMULATE: kind = 3; break;
                }
                if (kinds) kinds[i] = kind;
                if (ms && cudaEventElapsedTime(&ms[i], p->micro_ev[2 * i], p->micro_ev[2 * i + 1]) != cudaSuccess) return COSMA_B200_CUDA_ERROR;
                if (wire_bytes) wire_bytes[i] = wire;
            }
            return COSMA_B200_OK;
        }
        *n = static_cast<int>(ops.size());
        if (!p->time_gemms || p->ev.size() < 2 * ops.size()) return COSMA_B200_INVALID_ARG;
        for (size_t i = 0; i < ops.size() && static_cast<int>(i) < cap; ++i) {
            const auto& op = ops[i];
            if (kinds) kinds[i] = kind_of(op);
            if (ms && cudaEventElapsedTime(&ms[i], p->ev[2 * i], p->ev[2 * i + 1]) != cudaSuccess) return COSMA_B200_CUDA_ERROR;
            if (wire_bytes) wire_bytes[i] = wire_of(op);
        }
        return COSMA_B200_OK;
    });
}

/* The overlapped form of the plan's schedule (include/cosma/overlap.hpp): *enabled = 0 when the plan runs its ops serially (why: a
 * one-line reason). buf receives OverlapProgram::serialize(); est_ms = {serial, overlapped, communication} estimates of the planner. */
int cosma_b200_plan_overlap_export(void* plan, int64_t* buf, int64_t cap, int64_t* len, int* enabled, char* why, int why_len, double* est_ms) {
    return guarded("cosma_b200_plan_overlap_export", [&]() -> int {
        Plan* p = static_cast<Plan*>(plan);
        if (!p || !len || !enabled) return COSMA_B200_INVALID_ARG;
        *enabled = p->overlap.enabled ? 1 : 0;
        const auto v = p->overlap.serialize();
        *len = static_cast<int64_t>(v.size());
        if (buf && cap >= *len && !v.empty()) std::memcpy(buf, v.data(), v.size() * sizeof(int64_t));
        if (why && why_len > 0) {
            std::strncpy(why, p->overlap.why.c_str(), static_cast<size_t>(why_len) - 1);
            why[why_len - 1] = 0;
        }
        if (est_ms) { est_ms[0] = p->overlap.est_serial_ms; est_ms[1] = p->overlap.est_overlap_ms; est_ms[2] = p->overlap.est_comm_ms; }
        return COSMA_B200_OK;
    });
}

}  // extern "C"
