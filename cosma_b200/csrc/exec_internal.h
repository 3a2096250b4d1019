// Internal types shared by the executors (multiply_exec.cu, transform_exec.cu, layout_multiply.cu).
#pragma once
#include "../../include/cosma_b200.h"
#include "nccl_dyn.h"
#include "relayout_sm100.h"
#include "host_mirror.h"

#include <cosma/overlap.hpp>
#include <cosma/schedule.hpp>
#include <costa/transform_plan.hpp>

#include <cstdint>
#include <map>
#include <new>
#include <stdexcept>
#include <memory>
#include <string>
#include <vector>

namespace cosma_b200 {
void set_last_error(const std::string& msg);

struct LayoutMultiplyState;  // layout_multiply.cu

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, size = 1;
    // state cached across ?multiply_using_layout / p?gemm calls on this communicator (the reference caches the
    // communicator + strategy in its context, context.cpp:80-125); destroyed with the communicator
    std::map<std::string, LayoutMultiplyState*> layout_states;
    LayoutMultiplyState* last_layout_state = nullptr;
    unsigned long long layout_state_clock = 0;  // use counter for the least-recently-used bound on layout_states
    ~Comm();
};

// zero-SM peer transport of the overlapped schedules (peer_transport.h / peer_transport.cu)
struct PeerLink {
    int micro = -1;               // index of the micro-op (ALLGATHER or EXCHANGE) in Plan::overlap.ops
    char* landing[2] = {nullptr, nullptr};  // where to write in the mate's arena (mapped); [1]: the exchange's beta == 0 alternative
    uint32_t* mate_flags = nullptr;         // the mate's flag pair for this op (mapped): [0] ENTERED, [1] ARRIVED
    uint32_t* my_flags = nullptr;           // this rank's flag pair for this op
};

struct PeerTransport {
    bool ready = false;
    void* bound[3] = {nullptr, nullptr, nullptr};  // the arenas the landing zones were exchanged for
    char* flag_block = nullptr;   // own allocation (2 MiB: one IPC allocation of its own): flags at the front, epoch scratch behind
    std::vector<PeerLink> links;  // one per overlapped communication op, in program order
    std::vector<std::string> opened;  // IPC handles this plan holds a reference on (process-wide cache)
    uint32_t epoch = 0;
};

struct Plan {
    cosma::Schedule schedule;
    char dtype = 'd';
    int elem_reals = 1;    // real scalars per element: 1 (s, d) or 2 (c, z)
    int real_bytes = 8;    // 4 (s, c) or 8 (d, z)
    int elem_bytes() const { return elem_reals * real_bytes; }
    std::vector<ncclComm_t> ring_comms;  // by Schedule::rings() index
    int last_launches = 0;
    std::vector<float> gemm_ms;  // optional per-GEMM timing of the last run
    std::vector<cudaEvent_t> ev;
    bool time_gemms = false;
    // communication / computation overlap (host/overlap.cpp): the micro-op lowering of the schedule's tail. enabled only if EVERY rank of
    // the strategy lowers (ring mates must agree on the protocol); then the ring communicators are limited to `reserved` CTAs and
    // narrow GEMMs leave as many SMs free.
    cosma::OverlapProgram overlap;
    bool overlap_job = false;  // every active rank of the job lowers (the same value on every rank, idle ones included)
    int reserved = 0;
    bool rings_capped = false;  // the ring communicators were split with maxCTAs = reserved (NCCL transport of the overlapped program)
    cudaStream_t comm_stream = nullptr;
    std::vector<cudaEvent_t> micro_ev;  // [2 * micro-op] start / end, + 1 entry event
    bool last_run_overlapped = false;
    Comm* parent = nullptr;  // the communicator the plan was created on (verdicts that every rank must share)
    // copy-engine transports of the overlapped ops, one per bound set of arenas (cosma_b200_plan_bind_arenas: the caller's arenas, the
    // plan's own ones of the host-pointer entry point); a multiply on arenas that are not bound uses NCCL
    std::vector<std::unique_ptr<PeerTransport>> peers;
    // library-owned device arenas for the host-pointer entry point (allocated on first use)
    char* owned[3] = {nullptr, nullptr, nullptr};
    bool owned_bound = false;  // cosma_b200_plan_bind_arenas has been tried on them
    // column-panel pipelining of the host-pointer entry point (COSMA_B200_HOST_PANELS, multiply_exec.cu): the plan of one panel
    // (m, n / c, k, same strategy, this plan's ring communicators borrowed), two B / C arena sets, copy streams and events
    bool borrowed_comms = false;
    Plan* panel_plan = nullptr;
    int panel_count = 0;
    char* panel_arena[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [set][0 = B, 1 = C]
    cudaStream_t panel_in = nullptr, panel_out = nullptr;
    std::vector<cudaEvent_t> panel_ev;
};

// One piece of a rank's local B (or C) that belongs to column panel j: `len` elements from src_off of the local buffer correspond to
// dst_off of the panel's local buffer.
struct PanelPiece {
    std::int64_t src_off, len, dst_off;
};
// Column panel j of c of this rank's local matrices. The column ranges of all ranks' B blocks and C blocks are
// refined into elementary ranges (each inside one B range and one C range); panel j takes the j-th c-th of every elementary range:
// of local B those inside the rank's B columns, of local C those inside its C columns, each a contiguous piece of the column-major
// local buffer, concatenated in column order. False when the layout does not allow it (several blocks per rank, a width that c does
// not divide).
bool host_panel_pieces(const cosma::Schedule& schedule, int rank, int c, int j, std::vector<PanelPiece>& b_pieces,
                       std::vector<PanelPiece>& c_pieces);

// host copies of the local matrices for the one GEMM of a schedule that can stream them (multiply_exec.cu)
struct HostOperands {
    const void* A = nullptr;
    const void* B = nullptr;
    const void* C_in = nullptr;
    void* C_out = nullptr;
};
// skip_allgather_mask: bit x set = the allgathers of matrix x are skipped (the gathered copy in the arena is still valid)
int plan_run(Plan& plan, const double* alpha, const double* beta, void* A, void* B, void* C, cudaStream_t stream,
             const HostOperands* host = nullptr, unsigned skip_allgather_mask = 0);

// costa::transform on the device: pack kernel -> grouped ncclSend/ncclRecv -> unpack kernel
struct TransformPlan {
    costa::transform_plan host;
    char dtype = 'd';
    Comm* comm = nullptr;
    char* send_buf = nullptr;
    char* recv_buf = nullptr;
    RelayoutBatch stage1;  // pack + local pieces
    RelayoutBatch stage2;  // unpack pieces
    int last_launches = 0;
    bool materialised = false;  // device buffers allocated and piece lists uploaded (all of them)
    int failed = COSMA_B200_OK;  // status of a materialisation that failed: the plan is dead
    HostMirror mirror;          // device copies of the layout blocks that live in host memory (none: inactive)
    ~TransformPlan();
};
// Builds the device plan (allocates buffers, uploads piece lists). comm may be null (planning only / single rank).
int transform_plan_build(Comm* comm, int rank, int nranks, char dtype, const std::vector<costa::transform_spec>& specs,
                         std::unique_ptr<TransformPlan>& out);
int transform_plan_run(TransformPlan& plan, cudaStream_t stream);

#define COSMA_B200_NCCL_TRY(call)                                                                        \
    do {                                                                                                 \
        ncclResult_t r_ = (call);                                                                        \
        if (r_ != ncclSuccess) {                                                                         \
            ::cosma_b200::set_last_error(std::string(#call) + ": " + ::cosma_b200::nccl()->GetErrorString(r_)); \
            return COSMA_B200_NCCL_ERROR;                                                                \
        }                                                                                                \
    } while (0)
// No C++ exception may cross the C ABI: entry points whose body can throw (std::bad_alloc, a Mapper asked for a rank it does
// not have, ...) run inside guarded(), which turns the exception into a status + cosma_b200_last_error().
template <typename F>
int guarded(const char* what, F&& body) noexcept {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        try { set_last_error(std::string(what) + ": out of host memory"); } catch (...) {}
        return COSMA_B200_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        try { set_last_error(std::string(what) + ": " + e.what()); } catch (...) {}
        return COSMA_B200_INVALID_ARG;
    } catch (...) {
        return COSMA_B200_INTERNAL_ERROR;
    }
}

#define COSMA_B200_CUDA_TRY(call)                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            ::cosma_b200::set_last_error(std::string(#call) + ": " + cudaGetErrorString(e_));            \
            return COSMA_B200_CUDA_ERROR;                                                                \
        }                                                                                                \
    } while (0)

}  // namespace cosma_b200
