// costa::transform on the device (reference libs/COSTA/src/costa/grid2grid/transform.cpp:46-128 exchange_async and
// :231-282 transform): the reference packs on the host with OpenMP, posts MPI_Isend/Irecv per peer, copies local
// blocks, and unpacks each message as it arrives. Here the operands live in HBM:
//     stage 1 (one kernel)  : pack every outgoing piece into a per-peer segment of the send buffer, and move the
//                             pieces that stay on this rank straight into their target blocks (full transform)
//     exchange              : ONE NCCL group of ncclSend/ncclRecv over NVLink (the all-to-all-v)
//     stage 2 (one kernel)  : unpack, applying transpose / conjugate / alpha / beta
// Everything is queued on the caller's stream; plans (piece lists on the device, buffers) are reusable.
#include "exec_internal.h"

#include <cstring>

namespace cosma_b200 {

TransformPlan::~TransformPlan() {
    relayout_free(stage1);
    relayout_free(stage2);
    if (send_buf) cudaFree(send_buf);
    if (recv_buf) cudaFree(recv_buf);
}

int transform_plan_build(Comm* comm, int rank, int nranks, char dtype, const std::vector<costa::transform_spec>& specs,
                         std::unique_ptr<TransformPlan>& out) {
    const int eb = dtype_bytes(dtype);
    if (eb == 0) {
        set_last_error("transform: dtype must be one of s, d, c, z");
        return COSMA_B200_INVALID_ARG;
    }
    auto tp = std::make_unique<TransformPlan>();
    tp->dtype = dtype;
    tp->comm = comm;
    try {
        tp->host = costa::plan_transform(specs, rank, nranks, eb);
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
    // remember where the caller's blocks are: if they turn out to be host memory when the plan first runs, they are mirrored
    for (const auto& sp : specs) {
        const bool reads_target = sp.beta[0] != 0.0 || sp.beta[1] != 0.0;
        for (int side = 0; side < 2; ++side) {
            const costa::erased_layout* L = side == 0 ? sp.from : sp.to;
            if (!L) continue;
            for (const auto& b : L->blocks) {
                const size_t rows = L->grid.grid.rows_split[b.bi + 1] - L->grid.grid.rows_split[b.bi];
                const size_t cols = L->grid.grid.cols_split[b.bj + 1] - L->grid.grid.cols_split[b.bj];
                const size_t run = L->ordering == 'R' ? cols : rows, runs = L->ordering == 'R' ? rows : cols;
                tp->mirror.add(b.data, static_cast<size_t>(std::max<std::int64_t>(b.ld, static_cast<std::int64_t>(run))) * eb, run * eb, runs,
                               side == 1, reads_target);
            }
        }
    }
    out = std::move(tp);
    return COSMA_B200_OK;
}

namespace {
// device side of a plan: buffers + uploaded piece lists (separate from planning so tests can plan without a GPU)
int materialise_unchecked(TransformPlan& tp) {
    auto& h = tp.host;
    int ms = tp.mirror.build();
    if (ms != COSMA_B200_OK) return ms;
    if (tp.mirror.active()) {  // host-resident blocks: the pieces address their device mirrors from now on
        for (auto& p : h.pack) p.src = tp.mirror.translate(p.src);
        for (auto& p : h.local) { p.src = tp.mirror.translate(p.src); p.dst = tp.mirror.translate(p.dst); }
        for (auto& p : h.unpack) p.dst = tp.mirror.translate(p.dst);
    }
    if (h.total_send > 0 && cudaMalloc(reinterpret_cast<void**>(&tp.send_buf), h.total_send) != cudaSuccess) {
        set_last_error("transform: cudaMalloc of the send buffer failed");
        return COSMA_B200_OUT_OF_MEMORY;
    }
    if (h.total_recv > 0 && cudaMalloc(reinterpret_cast<void**>(&tp.recv_buf), h.total_recv) != cudaSuccess) {
        set_last_error("transform: cudaMalloc of the receive buffer failed");
        return COSMA_B200_OUT_OF_MEMORY;
    }
    RelayoutHostList l1, l2;
    relayout_normalise(h.pack, nullptr, tp.send_buf, h.elem_bytes, h.specs, l1);
    relayout_normalise(h.local, nullptr, nullptr, h.elem_bytes, h.specs, l1);
    int st = relayout_upload(l1, tp.stage1);
    if (st != COSMA_B200_OK) return st;
    relayout_normalise(h.unpack, tp.recv_buf, nullptr, h.elem_bytes, h.specs, l2);
    return relayout_upload(l2, tp.stage2);
}

// A plan counts as materialised only once EVERY allocation and upload has succeeded. A plan whose materialisation failed (out of
// device memory, ...) is dead: its partial device state is released, and every later run reports the failure again instead of
// launching empty batches (the piece lists were possibly already rewritten to mirror addresses, so a retry cannot be trusted);
// owners of plan caches drop such plans (layout_multiply.cu).
int materialise(TransformPlan& tp) {
    if (tp.materialised) return COSMA_B200_OK;
    if (tp.failed != COSMA_B200_OK) {
        set_last_error("transform: this plan could not be materialised on the device earlier (status " + std::to_string(tp.failed) + "); destroy it");
        return tp.failed;
    }
    const int st = materialise_unchecked(tp);
    if (st == COSMA_B200_OK) {
        tp.materialised = true;
        return st;
    }
    tp.failed = st;
    (void)cudaGetLastError();  // a failed cudaMalloc leaves a sticky-until-read error behind
    relayout_free(tp.stage1);
    relayout_free(tp.stage2);
    if (tp.send_buf) { cudaFree(tp.send_buf); tp.send_buf = nullptr; }
    if (tp.recv_buf) { cudaFree(tp.recv_buf); tp.recv_buf = nullptr; }
    return st;
}
}  // namespace

int transform_plan_run(TransformPlan& tp, cudaStream_t stream) {
    int st = materialise(tp);
    if (st != COSMA_B200_OK) return st;
    const auto& h = tp.host;
    tp.last_launches = 0;
    if (tp.mirror.active()) {
        st = tp.mirror.upload(stream);
        if (st != COSMA_B200_OK) return st;
    }
    st = relayout_launch(tp.stage1, tp.dtype, stream, &tp.last_launches);
    if (st != COSMA_B200_OK) return st;
    if (h.total_send > 0 || h.total_recv > 0) {
        if (!tp.comm || !tp.comm->comm) {
            set_last_error("transform: the plan exchanges data between ranks but was created without a communicator");
            return COSMA_B200_INVALID_ARG;
        }
        const NcclApi* N = nccl();
        COSMA_B200_NCCL_TRY(N->GroupStart());
        for (int p = 0; p < h.n_ranks; ++p) {
            if (h.send_bytes[p] > 0)
                COSMA_B200_NCCL_TRY(N->Send(tp.send_buf + h.send_off[p], static_cast<size_t>(h.send_bytes[p]), ncclChar, p, tp.comm->comm, stream));
            if (h.recv_bytes[p] > 0)
                COSMA_B200_NCCL_TRY(N->Recv(tp.recv_buf + h.recv_off[p], static_cast<size_t>(h.recv_bytes[p]), ncclChar, p, tp.comm->comm, stream));
        }
        COSMA_B200_NCCL_TRY(N->GroupEnd());
    }
    st = relayout_launch(tp.stage2, tp.dtype, stream, &tp.last_launches);
    if (st != COSMA_B200_OK) return st;
    return tp.mirror.active() ? tp.mirror.download(stream) : COSMA_B200_OK;
}

// erased_layout from the C struct (reference grid_from_clayout, src/cosma/cinterface.cpp:10-52)
costa::erased_layout layout_from_c(const cosma_b200_layout& l, char ordering, int nranks) {
    std::vector<int> br(l.nlocalblocks), bc(l.nlocalblocks);
    std::vector<void*> data(l.nlocalblocks);
    std::vector<std::int64_t> ld(l.nlocalblocks);
    for (int b = 0; b < l.nlocalblocks; ++b) {
        br[b] = l.localblocks[b].row;
        bc[b] = l.localblocks[b].col;
        data[b] = l.localblocks[b].data;
        ld[b] = l.localblocks[b].ld;
    }
    costa::erased_layout g = costa::erased_custom_layout(l.rowblocks, l.colblocks, l.rowsplit, l.colsplit, l.owners, l.nlocalblocks, br.data(),
                                                bc.data(), data.data(), ld.data(), ordering);
    g.grid.n_ranks = nranks;
    return g;
}

}  // namespace cosma_b200

using cosma_b200::Comm;
using cosma_b200::TransformPlan;
using cosma_b200::set_last_error;
using cosma_b200::guarded;

extern "C" {

int cosma_b200_transform_plan_create(void* comm, int rank, int nranks, char dtype, int n, const cosma_b200_layout* from,
                                     const cosma_b200_layout* to, const char* ordering_from, const char* ordering_to,
                                     const char* trans, const double* alpha, const double* beta, void** plan_out) {
    if (n < 0 || !plan_out || (n > 0 && (!from || !to))) return COSMA_B200_INVALID_ARG;
    Comm* c = static_cast<Comm*>(comm);
    if (c) { rank = c->rank; nranks = c->size; }
    try {
        std::vector<costa::erased_layout> F, T;
        F.reserve(n);
        T.reserve(n);
        for (int i = 0; i < n; ++i) {
            F.push_back(cosma_b200::layout_from_c(from[i], ordering_from ? ordering_from[i] : 'C', nranks));
            T.push_back(cosma_b200::layout_from_c(to[i], ordering_to ? ordering_to[i] : 'C', nranks));
        }
        std::vector<costa::transform_spec> specs(n);
        for (int i = 0; i < n; ++i) {
            specs[i].from = &F[i];
            specs[i].to = &T[i];
            specs[i].op = trans ? trans[i] : 'N';
            if (alpha) { specs[i].alpha[0] = alpha[2 * i]; specs[i].alpha[1] = alpha[2 * i + 1]; }
            if (beta) { specs[i].beta[0] = beta[2 * i]; specs[i].beta[1] = beta[2 * i + 1]; }
            if (dtype == 's' || dtype == 'd') specs[i].alpha[1] = specs[i].beta[1] = 0.0;
        }
        std::unique_ptr<TransformPlan> tp;
        const int st = cosma_b200::transform_plan_build(c, rank, nranks, dtype, specs, tp);
        if (st != COSMA_B200_OK) return st;
        *plan_out = tp.release();
        return COSMA_B200_OK;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
}

int cosma_b200_transform_run(void* plan, void* stream) {
    return guarded("cosma_b200_transform_run", [&]() -> int {
        if (!plan) return COSMA_B200_INVALID_ARG;
        return cosma_b200::transform_plan_run(*static_cast<TransformPlan*>(plan), static_cast<cudaStream_t>(stream));
    });
}

int cosma_b200_transform_plan_destroy(void* plan) {
    delete static_cast<TransformPlan*>(plan);
    return COSMA_B200_OK;
}

// Export format (int64):
//   [0] n_ranks  [1] elem_bytes  [2] total_send  [3] total_recv  [4] n_pack  [5] n_local  [6] n_unpack
//   then send_off[n_ranks], send_bytes[n_ranks], recv_off[n_ranks], recv_bytes[n_ranks]
//   then one 13-value record per piece, pack pieces first, then local, then unpack:
//     kind (0 pack, 1 local, 2 unpack), src, dst, src_ld, dst_ld, n_rows, n_cols, src_ordering, dst_ordering,
//     transpose, conjugate, transform, peer
//   (pack: dst is a byte offset into the send buffer; unpack: src is a byte offset into the receive buffer)
int cosma_b200_transform_plan_export(void* plan, int64_t* buf, int64_t cap, int64_t* len) {
    return guarded("cosma_b200_transform_plan_export", [&]() -> int {
        if (!plan || !len) return COSMA_B200_INVALID_ARG;
        const auto& h = static_cast<TransformPlan*>(plan)->host;
        std::vector<int64_t> v = {h.n_ranks, h.elem_bytes, h.total_send, h.total_recv, static_cast<int64_t>(h.pack.size()),
                                  static_cast<int64_t>(h.local.size()), static_cast<int64_t>(h.unpack.size())};
        for (const auto* arr : {&h.send_off, &h.send_bytes, &h.recv_off, &h.recv_bytes}) v.insert(v.end(), arr->begin(), arr->end());
        int kind = 0;
        for (const auto* list : {&h.pack, &h.local, &h.unpack}) {
            for (const auto& p : *list) {
                const int64_t rec[13] = {kind, reinterpret_cast<int64_t>(p.src), reinterpret_cast<int64_t>(p.dst), p.src_ld, p.dst_ld, p.n_rows,
                                         p.n_cols, p.src_ordering, p.dst_ordering, p.transpose, p.conjugate, p.transform, p.peer};
                v.insert(v.end(), rec, rec + 13);
            }
            ++kind;
        }
        *len = static_cast<int64_t>(v.size());
        if (buf && cap >= *len) std::memcpy(buf, v.data(), v.size() * sizeof(int64_t));
        return COSMA_B200_OK;
    });
}

int cosma_b200_transform_plan_stats(void* plan, int64_t* local_elements, int64_t* remote_elements, int* launches) {
    return guarded("cosma_b200_transform_plan_stats", [&]() -> int {
        if (!plan) return COSMA_B200_INVALID_ARG;
        const auto* tp = static_cast<TransformPlan*>(plan);
        if (local_elements) *local_elements = tp->host.local_elements;
        if (remote_elements) *remote_elements = tp->host.remote_elements;
        if (launches) *launches = tp->last_launches;
        return COSMA_B200_OK;
    });
}

int cosma_b200_relayout_batch(void* stream, char dtype, int n, const cosma_b200_piece* pieces) {
    const int eb = cosma_b200::dtype_bytes(dtype);
    if (eb == 0 || n < 0 || (n > 0 && !pieces)) return COSMA_B200_INVALID_ARG;
    if (n == 0) return COSMA_B200_OK;
    std::vector<costa::piece> ps(n);
    std::vector<costa::transform_spec> specs(n);
    for (int i = 0; i < n; ++i) {
        const auto& q = pieces[i];
        if ((q.src_ordering != 'C' && q.src_ordering != 'R') || (q.dst_ordering != 'C' && q.dst_ordering != 'R') || q.n_rows < 0 || q.n_cols < 0) {
            set_last_error("relayout_batch: bad ordering or negative size in piece " + std::to_string(i));
            return COSMA_B200_INVALID_ARG;
        }
        costa::piece& p = ps[i];
        p.src = q.src; p.dst = q.dst;
        p.n_rows = q.n_rows; p.n_cols = q.n_cols;
        p.src_ordering = q.src_ordering; p.dst_ordering = q.dst_ordering;
        p.transpose = q.transpose != 0; p.conjugate = q.conjugate != 0;
        // ld = 0: tight. op(src) is dr x dc; a tight block has ld = rows if 'C', cols if 'R'
        const int dr = p.transpose ? q.n_cols : q.n_rows, dc = p.transpose ? q.n_rows : q.n_cols;
        p.src_ld = q.src_ld > 0 ? q.src_ld : (p.src_ordering == 'C' ? q.n_rows : q.n_cols);
        p.dst_ld = q.dst_ld > 0 ? q.dst_ld : (p.dst_ordering == 'C' ? dr : dc);
        p.transform = i;
        specs[i].alpha[0] = q.alpha[0]; specs[i].alpha[1] = (dtype == 'c' || dtype == 'z') ? q.alpha[1] : 0.0;
        specs[i].beta[0] = q.beta[0]; specs[i].beta[1] = (dtype == 'c' || dtype == 'z') ? q.beta[1] : 0.0;
    }
    cosma_b200::RelayoutHostList list;
    cosma_b200::RelayoutBatch b;
    cosma_b200::relayout_normalise(ps, nullptr, nullptr, eb, specs, list);
    int st = cosma_b200::relayout_upload(list, b);
    if (st == COSMA_B200_OK) st = cosma_b200::relayout_launch(b, dtype, static_cast<cudaStream_t>(stream));
    // the piece list must outlive the kernel: free after the stream drains (this entry point is the convenience form;
    // plans keep their lists resident)
    if (!b.empty()) cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    cosma_b200::relayout_free(b);
    return st;
}

int cosma_b200_scalapack_layout(int lld, int mat_rows, int mat_cols, int ia, int ja, int sub_m, int sub_n, int mb, int nb, int nprow,
                                int npcol, char grid_order, int rsrc, int csrc, char data_ordering, int rank, int* rowblocks,
                                int* colblocks, int* rowsplit, int* colsplit, int* owners, int* nlocal, int* local_row, int* local_col,
                                int64_t* local_offset) {
    try {
        // element offsets: plan with a null base pointer and 1-byte elements
        const costa::erased_layout l = costa::erased_scalapack_layout(lld, mat_rows, mat_cols, ia, ja, sub_m, sub_n, mb, nb, nprow, npcol,
                                                                 grid_order, rsrc, csrc, nullptr, 1, data_ordering, rank);
        if (rowblocks) *rowblocks = l.grid.grid.n_rows();
        if (colblocks) *colblocks = l.grid.grid.n_cols();
        if (nlocal) *nlocal = static_cast<int>(l.blocks.size());
        if (rowsplit) std::copy(l.grid.grid.rows_split.begin(), l.grid.grid.rows_split.end(), rowsplit);
        if (colsplit) std::copy(l.grid.grid.cols_split.begin(), l.grid.grid.cols_split.end(), colsplit);
        if (owners) std::copy(l.grid.owners.begin(), l.grid.owners.end(), owners);
        for (size_t b = 0; b < l.blocks.size(); ++b) {
            if (local_row) local_row[b] = l.blocks[b].bi;
            if (local_col) local_col[b] = l.blocks[b].bj;
            if (local_offset) local_offset[b] = reinterpret_cast<int64_t>(l.blocks[b].data);
        }
        return COSMA_B200_OK;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
}

int cosma_b200_numroc(int n, int nb, int iproc, int isrcproc, int nprocs) { return costa::numroc(n, nb, iproc, isrcproc, nprocs); }

}  // extern "C"
