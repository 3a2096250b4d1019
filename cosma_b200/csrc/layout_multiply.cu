// multiply_using_layout and p?gemm on the device.
//
//   cosma::multiply_using_layout (reference src/cosma/multiply.cpp:78-213; C interface cinterface.cpp:54-150)
//   cosma::pxgemm               (reference src/cosma/cosma_pxgemm.cpp:16-388; ScaLAPACK symbols pxgemm.cpp:8-136)
//
// Same three phases as the reference: (1) COSTA-transform op(A), op(B) from the caller's layout into COSMA's native
// layout with alpha = 1, beta = 0; (2) cosma::multiply with alpha = 1, beta = 0; (3) COSTA-transform the COSMA-layout
// result into the caller's C with the caller's (alpha, beta). alpha, beta and op() are applied by the relayout kernels,
// never by the GEMM (multiply.cpp:186-204). Everything is resident in HBM and queued on one stream; the strategy, ring
// communicators, arenas and transform plans are cached per communicator (the reference caches communicator+strategy in
// its context, context.cpp:80-125, and re-derives the COSTA messages on every call).
#include "exec_internal.h"

#include <cosma/adapt_strategy.hpp>
#include <cosma/auto_strategy.hpp>
#include <cosma/environment_variables.hpp>
#include <costa/erased_layout.hpp>
#include <costa/grid2grid/comm_volume.hpp>

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace cosma_b200 {

costa::erased_layout layout_from_c(const cosma_b200_layout& l, char ordering, int nranks);  // transform_exec.cu

struct LayoutMultiplyState {
    void* plan = nullptr;  // cosma_b200 multiply plan (Plan*)
    char dtype = 'd';
    char* arena[3] = {nullptr, nullptr, nullptr};
    costa::erased_layout native[3];  // COSMA layouts of A, B, C with blocks pointing into the arenas
    struct Entry {
        std::string key;
        std::unique_ptr<TransformPlan> in, out;
        std::uint64_t stamp = 0;
    };
    std::vector<Entry> transforms;  // small LRU of transform plans keyed by the caller's layouts
    std::uint64_t clock = 0;
    int last_launches = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // start | A,B relayouted | multiplied | C relayouted
    std::int64_t last_elements[4] = {0, 0, 0, 0};              // in: local, remote; out: local, remote
    // rank relabelling (reference multiply.cpp:136-152): this rank plays COSMA rank perm[rank] on `relabelled`, the parent
    // communicator re-split with key perm[rank]; identity -> relabelled == nullptr and the plan runs on the parent
    std::vector<int> perm;
    Comm* relabelled = nullptr;
    std::uint64_t last_used = 0;    // per-communicator call counter at the last use (least recently used state is dropped first)
    std::uint64_t share_bytes = 0;  // (|A| + |B| + |C|) / ranks: the rank-independent size the cache bound counts
    ~LayoutMultiplyState() {
        transforms.clear();
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        for (auto& a : arena)
            if (a) cudaFree(a);
        if (plan) cosma_b200_plan_destroy(plan);
        if (relabelled) {
            if (relabelled->comm && nccl()) nccl()->CommDestroy(relabelled->comm);
            delete relabelled;
        }
    }
};

Comm::~Comm() {
    for (auto& kv : layout_states) delete kv.second;
}

namespace {

constexpr size_t kMaxCachedTransforms = 8;

// process grid of a ScaLAPACK-style call: what BLACS would answer for the context in desc[1]
// (reference blacs.hpp:5-35, scalapack.cpp:3-46)
struct Grid {
    Comm* comm = nullptr;
    char order = 'R';
    int nprow = 1, npcol = 1;
    // device staging for host-resident operands (N4), grown on demand
    char* stage[3] = {nullptr, nullptr, nullptr};
    size_t stage_bytes[3] = {0, 0, 0};
    // transform plans of p?tran / p?gemr2d keyed by descriptors, pointers, op and scalars (small LRU)
    struct Cached {
        std::string key;
        std::unique_ptr<TransformPlan> plan;
        std::uint64_t stamp = 0;
    };
    std::vector<Cached> transforms;
    std::uint64_t clock = 0;
    int last_launches = 0;
    ~Grid() {
        transforms.clear();
        for (auto& s : stage)
            if (s) cudaFree(s);
    }
};

void append(std::string& key, const void* p, size_t n) { key.append(static_cast<const char*>(p), n); }

void append_layout(std::string& key, const costa::erased_layout& l) {
    append(key, l.grid.grid.rows_split.data(), l.grid.grid.rows_split.size() * sizeof(int));
    append(key, l.grid.grid.cols_split.data(), l.grid.grid.cols_split.size() * sizeof(int));
    append(key, l.grid.owners.data(), l.grid.owners.size() * sizeof(int));
    for (const auto& b : l.blocks) {
        append(key, &b.bi, sizeof(int));
        append(key, &b.bj, sizeof(int));
        append(key, &b.data, sizeof(void*));
        append(key, &b.ld, sizeof(b.ld));
    }
    key.push_back(l.ordering);
    key.push_back('|');
}

bool is_zero(const double* v, bool cplx) { return v[0] == 0.0 && (!cplx || v[1] == 0.0); }

// C = beta * C on the caller's layout (grid_layout::scale_by, reference grid_layout.hpp:55-63); beta == 0 stores zeros
int scale_layout(char dtype, const costa::erased_layout& C, const double* beta, cudaStream_t stream) {
    const bool cplx = dtype == 'c' || dtype == 'z';
    if (beta[0] == 1.0 && (!cplx || beta[1] == 0.0)) return COSMA_B200_OK;
    std::vector<costa::piece> ps;
    const int eb = dtype_bytes(dtype);
    const bool beta_zero = is_zero(beta, cplx);
    HostMirror mirror;  // blocks in host memory are scaled through a device mirror
    for (const auto& b : C.blocks) {
        const size_t rows = C.grid.grid.rows_split[b.bi + 1] - C.grid.grid.rows_split[b.bi];
        const size_t cols = C.grid.grid.cols_split[b.bj + 1] - C.grid.grid.cols_split[b.bj];
        const size_t run = C.ordering == 'R' ? cols : rows, runs = C.ordering == 'R' ? rows : cols;
        mirror.add(b.data, static_cast<size_t>(std::max<std::int64_t>(b.ld, static_cast<std::int64_t>(run))) * eb, run * eb, runs, true, !beta_zero);
    }
    int ms = mirror.build();
    if (ms != COSMA_B200_OK) return ms;
    if (mirror.active() && (ms = mirror.upload(stream)) != COSMA_B200_OK) return ms;
    for (const auto& b : C.blocks) {
        costa::piece p;
        p.n_rows = C.grid.grid.rows_split[b.bi + 1] - C.grid.grid.rows_split[b.bi];
        p.n_cols = C.grid.grid.cols_split[b.bj + 1] - C.grid.grid.cols_split[b.bj];
        p.src = mirror.translate(b.data); p.dst = mirror.translate(b.data);
        p.src_ld = p.dst_ld = b.ld;
        p.src_ordering = p.dst_ordering = C.ordering;
        p.scale_only = true;
        p.transform = 0;
        ps.push_back(p);
    }
    std::vector<costa::transform_spec> specs(1);
    specs[0].alpha[0] = 0.0; specs[0].alpha[1] = 0.0;
    specs[0].beta[0] = beta[0]; specs[0].beta[1] = cplx ? beta[1] : 0.0;
    RelayoutHostList list;
    RelayoutBatch batch;
    relayout_normalise(ps, nullptr, nullptr, dtype_bytes(dtype), specs, list);
    int st = relayout_upload(list, batch);
    if (st == COSMA_B200_OK) st = relayout_launch(batch, dtype, stream);
    if (st == COSMA_B200_OK && mirror.active()) st = mirror.download(stream);
    if (!batch.empty() || mirror.active()) cudaStreamSynchronize(stream);
    relayout_free(batch);
    return st;
}

// COSMA's native layouts of A, B, C as assigned grids with owners in COSMA rank labels (Mapper::get_layout_grid, reference
// mapper.cpp:369-414): global information, the same on every rank, cached per problem
struct NativeGrids {
    costa::assigned_grid2D g[3];
};
const NativeGrids& native_grids(int nranks, int m, int n, int k, const char* steps, char dtype) {
    static std::mutex mu;
    static std::map<std::string, NativeGrids> cache;
    const std::string key = std::string(1, dtype) + std::to_string(nranks) + ":" + std::to_string(m) + ":" + std::to_string(n) + ":" + std::to_string(k) + ":" + (steps ? steps : "");
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    NativeGrids ng;
    const cosma::Strategy strategy = cosma::automatic_strategy(m, n, k, static_cast<size_t>(nranks), steps ? steps : "", dtype_bytes(dtype));  // as the plan
    for (int x = 0; x < 3; ++x) {
        const cosma::Mapper mapper("ABC"[x], strategy, 0);
        ng.g[x].grid.rows_split = mapper.row_split();
        ng.g[x].grid.cols_split = mapper.col_split();
        ng.g[x].n_ranks = nranks;
        const auto owners = mapper.grid_owners();
        const int nr = ng.g[x].grid.n_rows(), nc = ng.g[x].grid.n_cols();
        ng.g[x].owners.resize(static_cast<size_t>(nr) * nc);
        for (int i = 0; i < nr; ++i)
            for (int j = 0; j < nc; ++j) ng.g[x].owners[static_cast<size_t>(i) * nc + j] = owners[i][j];
    }
    return cache.emplace(key, std::move(ng)).first->second;
}

bool relabelling_enabled() {
    // as the reference, which always relabels (multiply.cpp:136-152); COSMA_B200_REORDER_RANKS=OFF keeps the caller's labels
    static const bool on = cosma::get_bool_env_var("COSMA_B200_REORDER_RANKS", true);
    return on;
}

// how many problems (dtype, m, n, k, strategy, relabelling) keep their plan and arenas per communicator
size_t max_cached_problems() {
    static const size_t n = [] {
        const char* v = std::getenv("COSMA_B200_CACHED_PROBLEMS");
        const long long x = v && *v ? std::atoll(v) : 64;
        return static_cast<size_t>(x < 1 ? 1 : x);
    }();
    return n;
}
// ... and how many bytes of matrices per rank they may stand for together (COSMA_B200_CACHED_PROBLEMS_MB, default 32 GiB of 180 GB)
std::uint64_t max_cached_bytes() {
    static const std::uint64_t n = [] {
        const char* v = std::getenv("COSMA_B200_CACHED_PROBLEMS_MB");
        const long long x = v && *v ? std::atoll(v) : 32768;
        return static_cast<std::uint64_t>(x < 0 ? 0 : x) << 20;
    }();
    return n;
}

// perm: this rank plays COSMA rank perm[rank] (an involution; empty or identity = no relabelling)
int get_state(Comm* c, char dtype, int m, int n, int k, const char* steps, const std::vector<int>& perm, LayoutMultiplyState** out) {
    std::string key = std::string(1, dtype) + ":" + std::to_string(m) + ":" + std::to_string(n) + ":" + std::to_string(k) + ":" + (steps ? steps : "");
    bool relabel = false;
    for (size_t r = 0; r < perm.size(); ++r) relabel = relabel || perm[r] != static_cast<int>(r);
    if (relabel) {
        key += ":perm";
        for (int v : perm) key += "," + std::to_string(v);
    }
    auto it = c->layout_states.find(key);
    if (it != c->layout_states.end()) {
        it->second->last_used = ++c->layout_state_clock;
        *out = it->second;
        return COSMA_B200_OK;
    }
    // bounded cache: every state owns three device arenas and a set of ring communicators, so an application that walks through many
    // shapes (the reference keeps ONE strategy per context and replans, context.cpp:80-125) must not accumulate them. Bounds: a
    // number of problems and a number of bytes, the latter counted as the three matrices' share per rank -- a figure that is the
    // same on every rank; every rank makes the same calls in the same order, so every rank drops the same states.
    const std::uint64_t share = static_cast<std::uint64_t>((static_cast<double>(m) * k + static_cast<double>(k) * n + static_cast<double>(m) * n) *
                                                           dtype_bytes(dtype) / std::max(c->size, 1));
    for (;;) {
        std::uint64_t held = 0;
        for (const auto& kv : c->layout_states) held += kv.second->share_bytes;
        if (c->layout_states.empty() || (c->layout_states.size() < max_cached_problems() && held + share <= max_cached_bytes())) break;
        auto oldest = c->layout_states.begin();
        for (auto jt = c->layout_states.begin(); jt != c->layout_states.end(); ++jt)
            if (jt->second->last_used < oldest->second->last_used) oldest = jt;
        COSMA_B200_CUDA_TRY(cudaDeviceSynchronize());  // its last multiply may still be queued
        if (c->last_layout_state == oldest->second) c->last_layout_state = nullptr;
        delete oldest->second;
        c->layout_states.erase(oldest);
    }
    auto st = std::make_unique<LayoutMultiplyState>();
    st->dtype = dtype;
    st->last_used = ++c->layout_state_clock;
    st->share_bytes = share;
    Comm* pc = c;  // the communicator the multiply runs on
    if (relabel) {
        st->perm = perm;
        ncclComm_t sub = nullptr;
        COSMA_B200_NCCL_TRY(nccl()->CommSplit(c->comm, 0, perm[c->rank], &sub, nullptr));
        st->relabelled = new Comm;
        st->relabelled->comm = sub;
        st->relabelled->rank = perm[c->rank];
        st->relabelled->size = c->size;
        pc = st->relabelled;
    }
    int rc = cosma_b200_plan_create(pc, pc->rank, pc->size, m, n, k, steps ? steps : "", dtype, &st->plan);
    if (rc != COSMA_B200_OK) return rc;
    Plan* plan = static_cast<Plan*>(st->plan);
    const int eb = dtype_bytes(dtype);
    const int P = static_cast<int>(plan->schedule.strategy().P);
    for (int x = 0; x < 3; ++x) {
        const size_t bytes = std::max<std::int64_t>(plan->schedule.arena_elements(x), 1) * eb;
        if (cudaMalloc(reinterpret_cast<void**>(&st->arena[x]), bytes) != cudaSuccess) {
            set_last_error("multiply_using_layout: cudaMalloc of a COSMA arena failed");
            return COSMA_B200_OUT_OF_MEMORY;
        }
        // the COSMA layout as a COSTA grid (reference Mapper::get_layout_grid, mapper.cpp:369-414, and
        // CosmaMatrix::get_grid_layout, matrix.cpp:393-429: one column-major block per Mapper block, ld = rows)
        const cosma::Mapper& mapper = plan->schedule.mapper(x);
        costa::erased_layout& L = st->native[x];
        L.ordering = 'C';
        L.grid.grid.rows_split = mapper.row_split();
        L.grid.grid.cols_split = mapper.col_split();
        L.grid.n_ranks = c->size;
        const auto owners = mapper.grid_owners();
        const int nr = L.grid.grid.n_rows(), nc = L.grid.grid.n_cols();
        L.grid.owners.resize(static_cast<size_t>(nr) * nc);
        // owners in the PARENT communicator's labels: COSMA rank q is played by parent rank perm[q] (perm is an involution)
        for (int i = 0; i < nr; ++i)
            for (int j = 0; j < nc; ++j) L.grid.owners[static_cast<size_t>(i) * nc + j] = relabel ? perm[owners[i][j]] : owners[i][j];
        if (pc->rank < P) {
            const auto& blocks = mapper.initial_layout(pc->rank);
            const auto& offs = mapper.blocks_offsets(pc->rank);
            for (size_t b = 0; b < blocks.size(); ++b) {
                const auto& rs = L.grid.grid.rows_split;
                const auto& cs = L.grid.grid.cols_split;
                const int bi = static_cast<int>(std::upper_bound(rs.begin(), rs.end(), blocks[b].rows.first()) - rs.begin()) - 1;
                const int bj = static_cast<int>(std::upper_bound(cs.begin(), cs.end(), blocks[b].cols.first()) - cs.begin()) - 1;
                L.blocks.push_back(costa::local_block{bi, bj, st->arena[x] + static_cast<std::int64_t>(offs[b]) * eb,
                                                      static_cast<std::int64_t>(blocks[b].rows.length())});
            }
        }
    }
    // the arenas never change: overlapped transfers of the multiply go through copy engines (collective; no-op for other plans)
    rc = cosma_b200_plan_bind_arenas(st->plan, st->arena[0], st->arena[1], st->arena[2], nullptr);
    if (rc != COSMA_B200_OK) return rc;
    *out = st.get();
    c->layout_states[key] = st.release();
    return COSMA_B200_OK;
}

int layout_multiply(Comm* c, char dtype, char ta, char tb, int m, int n, int k, const double* alpha, const double* beta,
                    const costa::erased_layout& A, const costa::erased_layout& B, const costa::erased_layout& C, const char* steps,
                    cudaStream_t stream, int* launches) {
    const bool cplx = dtype == 'c' || dtype == 'z';
    if (launches) *launches = 0;
    // corner cases allowed by the BLAS standard (multiply.cpp:96-109, cosma_pxgemm.cpp:36-46)
    if (m == 0 || n == 0) return COSMA_B200_OK;
    if (k == 0 || is_zero(alpha, cplx)) return scale_layout(dtype, C, beta, stream);
    // rank relabelling (reference multiply.cpp:136-152, cosma_pxgemm.cpp:255-271): relabel the COSMA ranks so
    // that as much of A, B and C as possible is already where COSMA's layout wants it; every rank derives the same
    // permutation from the global grids
    std::vector<int> perm;
    const bool adapted = steps && *steps;  // a strategy adapted to the caller's grid is not relabelled (cosma_pxgemm.cpp:255-283)
    if (c->size > 1 && c->comm && relabelling_enabled() && !adapted) {
        const NativeGrids& ng = native_grids(c->size, m, n, k, steps, dtype);
        costa::comm_volume vol = costa::communication_volume(A.grid, ng.g[0], ta);
        vol += costa::communication_volume(B.grid, ng.g[1], tb);
        vol += costa::communication_volume(ng.g[2], C.grid, 'N');
        bool reordered = false;
        perm = costa::optimal_reordering(vol, c->size, reordered);
        if (!reordered) perm.clear();
    }
    LayoutMultiplyState* st = nullptr;
    int rc = get_state(c, dtype, m, n, k, steps, perm, &st);
    if (rc != COSMA_B200_OK) return rc;

    std::string key;
    key.push_back(ta); key.push_back(tb);
    append(key, alpha, 2 * sizeof(double));
    append(key, beta, 2 * sizeof(double));
    append_layout(key, A);
    append_layout(key, B);
    append_layout(key, C);
    LayoutMultiplyState::Entry* entry = nullptr;
    for (auto& e : st->transforms)
        if (e.key == key) entry = &e;
    if (!entry) {
        if (st->transforms.size() >= kMaxCachedTransforms) {
            size_t oldest = 0;
            for (size_t i = 1; i < st->transforms.size(); ++i)
                if (st->transforms[i].stamp < st->transforms[oldest].stamp) oldest = i;
            // the evicted plans may still be running on the stream
            cudaStreamSynchronize(stream);
            st->transforms.erase(st->transforms.begin() + oldest);
        }
        LayoutMultiplyState::Entry e;
        e.key = key;
        std::vector<costa::transform_spec> in(2), out(1);
        in[0].from = &A; in[0].to = &st->native[0]; in[0].op = ta;
        in[1].from = &B; in[1].to = &st->native[1]; in[1].op = tb;
        out[0].from = &st->native[2]; out[0].to = &C; out[0].op = 'N';
        out[0].alpha[0] = alpha[0]; out[0].alpha[1] = cplx ? alpha[1] : 0.0;
        out[0].beta[0] = beta[0]; out[0].beta[1] = cplx ? beta[1] : 0.0;
        rc = transform_plan_build(c, c->rank, c->size, dtype, in, e.in);
        if (rc != COSMA_B200_OK) return rc;
        rc = transform_plan_build(c, c->rank, c->size, dtype, out, e.out);
        if (rc != COSMA_B200_OK) return rc;
        st->transforms.push_back(std::move(e));
        entry = &st->transforms.back();
    }
    entry->stamp = ++st->clock;
    for (auto& e : st->ev)
        if (!e) COSMA_B200_CUDA_TRY(cudaEventCreate(&e));
    c->last_layout_state = st;
    st->last_elements[0] = entry->in->host.local_elements;
    st->last_elements[1] = entry->in->host.remote_elements;
    st->last_elements[2] = entry->out->host.local_elements;
    st->last_elements[3] = entry->out->host.remote_elements;

    // a transform plan that fails to run (typically: out of device memory while materialising) must not stay in the cache, or the
    // next call with the same layouts would find it and skip the work
    auto drop_entry = [&](int status) {
        cudaStreamSynchronize(stream);
        for (size_t i = 0; i < st->transforms.size(); ++i)
            if (&st->transforms[i] == entry) {
                st->transforms.erase(st->transforms.begin() + i);
                break;
            }
        return status;
    };
    COSMA_B200_CUDA_TRY(cudaEventRecord(st->ev[0], stream));
    rc = transform_plan_run(*entry->in, stream);
    if (rc != COSMA_B200_OK) return drop_entry(rc);
    COSMA_B200_CUDA_TRY(cudaEventRecord(st->ev[1], stream));
    const double one[2] = {1.0, 0.0}, zero[2] = {0.0, 0.0};
    rc = cosma_b200_multiply(st->plan, one, zero, st->arena[0], st->arena[1], st->arena[2], stream);
    if (rc != COSMA_B200_OK) return rc;
    COSMA_B200_CUDA_TRY(cudaEventRecord(st->ev[2], stream));
    rc = transform_plan_run(*entry->out, stream);
    if (rc != COSMA_B200_OK) return drop_entry(rc);
    COSMA_B200_CUDA_TRY(cudaEventRecord(st->ev[3], stream));
    st->last_launches = entry->in->last_launches + entry->out->last_launches + cosma_b200_plan_last_launches(st->plan);
    if (launches) *launches = st->last_launches;
    return COSMA_B200_OK;
}


// D2H of a staged local array: only the rank's local rows of every column (the padding rows between the local row count and
// lld belong to the caller and were possibly never uploaded)
cudaError_t download_local(void* host, const void* dev, int lld, int loc_rows, int loc_cols, int eb, cudaStream_t stream) {
    if (loc_rows <= 0 || loc_cols <= 0) return cudaSuccess;
    return cudaMemcpy2DAsync(host, static_cast<size_t>(lld) * eb, dev, static_cast<size_t>(lld) * eb, static_cast<size_t>(loc_rows) * eb,
                             loc_cols, cudaMemcpyDeviceToHost, stream);
}

}  // namespace
}  // namespace cosma_b200

using namespace cosma_b200;

extern "C" {

static int xmultiply_using_layout(void* comm, char dtype, const char* transa, const char* transb, const double* alpha,
                                  const cosma_b200_layout* A, const cosma_b200_layout* B, const double* beta,
                                  const cosma_b200_layout* C, void* stream) {
    Comm* c = static_cast<Comm*>(comm);
    if (!c || !transa || !transb || !alpha || !beta || !A || !B || !C) return COSMA_B200_INVALID_ARG;
    try {
        const char ta = std::toupper(*transa), tb = std::toupper(*transb);
        const costa::erased_layout LA = layout_from_c(*A, 'C', c->size), LB = layout_from_c(*B, 'C', c->size),
                                 LC = layout_from_c(*C, 'C', c->size);
        const int m = LC.num_rows(), n = LC.num_cols();
        const int k = ta == 'N' ? LA.num_cols() : LA.num_rows();
        const double a2[2] = {alpha[0], (dtype == 'c' || dtype == 'z') ? alpha[1] : 0.0};
        const double b2[2] = {beta[0], (dtype == 'c' || dtype == 'z') ? beta[1] : 0.0};
        return layout_multiply(c, dtype, ta, tb, m, n, k, a2, b2, LA, LB, LC, "", static_cast<cudaStream_t>(stream), nullptr);
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
}

int cosma_b200_dmultiply_using_layout(void* comm, const char* transa, const char* transb, const double* alpha,
                                      const cosma_b200_layout* A, const cosma_b200_layout* B, const double* beta,
                                      const cosma_b200_layout* C, void* stream) {
    return xmultiply_using_layout(comm, 'd', transa, transb, alpha, A, B, beta, C, stream);
}
int cosma_b200_zmultiply_using_layout(void* comm, const char* transa, const char* transb, const double* alpha,
                                      const cosma_b200_layout* A, const cosma_b200_layout* B, const double* beta,
                                      const cosma_b200_layout* C, void* stream) {
    return xmultiply_using_layout(comm, 'z', transa, transb, alpha, A, B, beta, C, stream);
}

int cosma_b200_last_layout_multiply_stats(void* comm, float* ms3, int64_t* elements4, char* strategy, int strategy_len, int* launches) {
    return guarded("cosma_b200_last_layout_multiply_stats", [&]() -> int {
        Comm* c = static_cast<Comm*>(comm);
        if (!c || !c->last_layout_state) return COSMA_B200_INVALID_ARG;
        LayoutMultiplyState* st = c->last_layout_state;
        if (ms3)
            for (int i = 0; i < 3; ++i)
                if (cudaEventElapsedTime(&ms3[i], st->ev[i], st->ev[i + 1]) != cudaSuccess) return COSMA_B200_CUDA_ERROR;
        if (elements4) std::memcpy(elements4, st->last_elements, sizeof(st->last_elements));
        if (strategy) cosma_b200_plan_strategy(st->plan, strategy, strategy_len, nullptr);
        if (launches) *launches = st->last_launches;
        return COSMA_B200_OK;
    });
}

int cosma_b200_smultiply_using_layout(void* comm, const char* transa, const char* transb, const double* alpha,
                                      const cosma_b200_layout* A, const cosma_b200_layout* B, const double* beta,
                                      const cosma_b200_layout* C, void* stream) {
    return xmultiply_using_layout(comm, 's', transa, transb, alpha, A, B, beta, C, stream);
}
int cosma_b200_cmultiply_using_layout(void* comm, const char* transa, const char* transb, const double* alpha,
                                      const cosma_b200_layout* A, const cosma_b200_layout* B, const double* beta,
                                      const cosma_b200_layout* C, void* stream) {
    return xmultiply_using_layout(comm, 'c', transa, transb, alpha, A, B, beta, C, stream);
}

int cosma_b200_grid_create(void* comm, char order, int nprow, int npcol, void** grid_out) {
    return guarded("cosma_b200_grid_create", [&]() -> int {
        Comm* c = static_cast<Comm*>(comm);
        if (!c || !grid_out || nprow < 1 || npcol < 1 || nprow * npcol > c->size) {
            set_last_error("grid_create: need a communicator and nprow*npcol <= its size");
            return COSMA_B200_INVALID_ARG;
        }
        order = std::toupper(order);
        if (order != 'R' && order != 'C') return COSMA_B200_INVALID_ARG;
        auto* g = new Grid;
        g->comm = c;
        g->order = order;
        g->nprow = nprow;
        g->npcol = npcol;
        *grid_out = g;
        return COSMA_B200_OK;
    });
}
int cosma_b200_grid_destroy(void* grid) {
    delete static_cast<Grid*>(grid);
    return COSMA_B200_OK;
}
int cosma_b200_grid_info(void* grid, int* nprow, int* npcol, int* myrow, int* mycol) {
    return guarded("cosma_b200_grid_info", [&]() -> int {
        Grid* g = static_cast<Grid*>(grid);
        if (!g) return COSMA_B200_INVALID_ARG;
        if (nprow) *nprow = g->nprow;
        if (npcol) *npcol = g->npcol;
        int r = -1, cc = -1;
        if (g->comm->rank < g->nprow * g->npcol) costa::rank_to_grid(g->comm->rank, g->nprow, g->npcol, g->order, &r, &cc);
        if (myrow) *myrow = r;
        if (mycol) *mycol = cc;
        return COSMA_B200_OK;
    });
}

// descriptor fields (reference scalapack.hpp:11-47): [2] M, [3] N, [4] MB, [5] NB, [6] RSRC, [7] CSRC, [8] LLD
static int xpgemm(void* grid, char dtype, char transa, char transb, int m, int n, int k, const double* alpha, const void* a, int ia,
                  int ja, const int* desca, const void* b, int ib, int jb, const int* descb, const double* beta, void* c, int ic, int jc,
                  const int* descc, void* stream_v) {
    Grid* g = static_cast<Grid*>(grid);
    if (!g || !alpha || !beta || !desca || !descb || !descc) return COSMA_B200_INVALID_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    try {
        const bool cplx = dtype == 'c' || dtype == 'z';
        const int eb = dtype_bytes(dtype);
        if (eb == 0) return COSMA_B200_INVALID_ARG;
        const char ta = std::toupper(transa), tb = std::toupper(transb);
        if (m == 0 || n == 0) return COSMA_B200_OK;
        const int a_subm = ta == 'N' ? m : k, a_subn = ta == 'N' ? k : m;
        const int b_subm = tb == 'N' ? k : n, b_subn = tb == 'N' ? n : k;
        const int rank = g->comm->rank;
        const bool in_grid = rank < g->nprow * g->npcol;
        int myrow = 0, mycol = 0;
        if (in_grid) costa::rank_to_grid(rank, g->nprow, g->npcol, g->order, &myrow, &mycol);

        // host-resident operands (what a ScaLAPACK application passes): stage the rank's local arrays through HBM
        const void* user[3] = {a, b, c};
        const int* desc[3] = {desca, descb, descc};
        void* dev[3] = {const_cast<void*>(a), const_cast<void*>(b), c};
        bool staged[3] = {false, false, false};
        size_t local_bytes[3] = {0, 0, 0};
        const bool scale_only = k == 0 || is_zero(alpha, cplx);
        const bool c_whole = ic == 1 && jc == 1 && m == descc[2] && n == descc[3];
        for (int x = 0; x < 3; ++x) {
            if (!in_grid || !user[x]) continue;
            if (scale_only && x < 2) continue;
            if (!is_host_pointer(user[x])) continue;
            const int loc_cols = costa::numroc(desc[x][3], desc[x][5], mycol, desc[x][7], g->npcol);
            local_bytes[x] = static_cast<size_t>(desc[x][8]) * std::max(loc_cols, 0) * eb;
            if (local_bytes[x] == 0) continue;
            if (g->stage_bytes[x] < local_bytes[x]) {
                if (g->stage[x]) { cudaStreamSynchronize(stream); cudaFree(g->stage[x]); g->stage[x] = nullptr; g->stage_bytes[x] = 0; }
                if (cudaMalloc(reinterpret_cast<void**>(&g->stage[x]), local_bytes[x]) != cudaSuccess) {
                    set_last_error("p?gemm: cudaMalloc of a staging buffer failed");
                    return COSMA_B200_OUT_OF_MEMORY;
                }
                g->stage_bytes[x] = local_bytes[x];
            }
            staged[x] = true;
            dev[x] = g->stage[x];
            // C is uploaded unless it is overwritten completely (beta == 0 on the whole matrix)
            const bool upload = x < 2 || !(is_zero(beta, cplx) && c_whole);
            if (upload) COSMA_B200_CUDA_TRY(cudaMemcpyAsync(dev[x], user[x], local_bytes[x], cudaMemcpyHostToDevice, stream));
        }

        auto layout_of = [&](int x, int i0, int j0, int sm, int sn) {
            return costa::erased_scalapack_layout(desc[x][8], desc[x][2], desc[x][3], i0, j0, sm, sn, desc[x][4], desc[x][5], g->nprow, g->npcol,
                                               g->order, desc[x][6], desc[x][7], dev[x], eb, 'C', in_grid ? rank : -1);
        };
        const double a2[2] = {alpha[0], cplx ? alpha[1] : 0.0}, b2[2] = {beta[0], cplx ? beta[1] : 0.0};
        int rc;
        if (scale_only) {
            costa::erased_layout LC = layout_of(2, ic, jc, m, n);
            LC.grid.n_ranks = g->comm->size;
            rc = scale_layout(dtype, LC, b2, stream);
        } else {
            costa::erased_layout LA = layout_of(0, ia, ja, a_subm, a_subn), LB = layout_of(1, ib, jb, b_subm, b_subn),
                               LC = layout_of(2, ic, jc, m, n);
            LA.grid.n_ranks = LB.grid.n_ranks = LC.grid.n_ranks = g->comm->size;
            // COSMA_ADAPT_STRATEGY=ON (the reference's default; opt-in here): start the strategy with the steps that
            // reproduce the block-cyclic grid of the largest operand, so that this operand needs (almost) no relayout
            std::string steps;
            if (g->comm->size > 1 && cosma::env_var_defined("COSMA_ADAPT_STRATEGY") && cosma::get_adapt_strategy()) {
                auto of = [](const int* d, int i, int j) {
                    cosma::block_cyclic_desc b;
                    b.rows = d[2]; b.cols = d[3]; b.block_rows = d[4]; b.block_cols = d[5]; b.i = i; b.j = j;
                    return b;
                };
                const std::string prefix = cosma::adapt_strategy_to_block_cyclic_grid(m, n, k, g->comm->size, of(desca, ia, ja), of(descb, ib, jb),
                                                                                      of(descc, ic, jc), ta, tb, g->nprow, g->npcol, g->order);
                if (!prefix.empty()) steps = cosma::parse_strategy(m, n, k, static_cast<size_t>(g->comm->size), prefix).to_string();
            }
            rc = layout_multiply(g->comm, dtype, ta, tb, m, n, k, a2, b2, LA, LB, LC, steps.c_str(), stream, nullptr);
        }
        if (rc != COSMA_B200_OK) return rc;
        if (staged[2])
            COSMA_B200_CUDA_TRY(download_local(c, dev[2], descc[8], costa::numroc(descc[2], descc[4], myrow, descc[6], g->nprow),
                                               costa::numroc(descc[3], descc[5], mycol, descc[7], g->npcol), eb, stream));
        return COSMA_B200_OK;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
}

int cosma_b200_pdgemm(void* grid, char transa, char transb, int m, int n, int k, const double* alpha, const double* a, int ia, int ja,
                      const int* desca, const double* b, int ib, int jb, const int* descb, const double* beta, double* c, int ic, int jc,
                      const int* descc, void* stream) {
    return xpgemm(grid, 'd', transa, transb, m, n, k, alpha, a, ia, ja, desca, b, ib, jb, descb, beta, c, ic, jc, descc, stream);
}
int cosma_b200_pzgemm(void* grid, char transa, char transb, int m, int n, int k, const double* alpha, const double* a, int ia, int ja,
                      const int* desca, const double* b, int ib, int jb, const int* descb, const double* beta, double* c, int ic, int jc,
                      const int* descc, void* stream) {
    return xpgemm(grid, 'z', transa, transb, m, n, k, alpha, a, ia, ja, desca, b, ib, jb, descb, beta, c, ic, jc, descc, stream);
}

int cosma_b200_psgemm(void* grid, char transa, char transb, int m, int n, int k, const double* alpha, const float* a, int ia, int ja,
                      const int* desca, const float* b, int ib, int jb, const int* descb, const double* beta, float* c, int ic, int jc,
                      const int* descc, void* stream) {
    return xpgemm(grid, 's', transa, transb, m, n, k, alpha, a, ia, ja, desca, b, ib, jb, descb, beta, c, ic, jc, descc, stream);
}
int cosma_b200_pcgemm(void* grid, char transa, char transb, int m, int n, int k, const double* alpha, const float* a, int ia, int ja,
                      const int* desca, const float* b, int ib, int jb, const int* descb, const double* beta, float* c, int ic, int jc,
                      const int* descc, void* stream) {
    return xpgemm(grid, 'c', transa, transb, m, n, k, alpha, a, ia, ja, desca, b, ib, jb, descb, beta, c, ic, jc, descc, stream);
}

// ---- p?tran / p?tranu / p?tranc and p?gemr2d ------------------------------------------------------
// Reference: libs/COSTA/src/costa/pxtran_op/costa_pxtran_op.cpp:14-172 (sub(C) = beta*sub(C) + alpha*op(sub(A)), sub(C) is
// m x n and sub(A) n x m) and pxgemr2d/costa_pxgemr2d.cpp:14-168 (sub(C) = sub(A) between two process grids). Both are one
// costa::transform between two block-cyclic layouts: the same plan + R3/R4 kernels + NCCL exchange as p?gemm's relayouts.
static int xptransform(Grid* ga, Grid* gc, char dtype, char op, int m, int n, const double* alpha, const void* a, int ia, int ja,
                       const int* desca, const double* beta, void* c, int ic, int jc, const int* descc, cudaStream_t stream) {
    if (!ga || !gc || !desca || !descc || !alpha || !beta || ga->comm != gc->comm) return COSMA_B200_INVALID_ARG;
    try {
        const bool cplx = dtype == 'c' || dtype == 'z';
        const int eb = dtype_bytes(dtype);
        if (eb == 0) return COSMA_B200_INVALID_ARG;
        op = std::toupper(op);
        if (op != 'N' && op != 'T' && op != 'C') return COSMA_B200_INVALID_ARG;
        if (m == 0 || n == 0) return COSMA_B200_OK;
        Comm* comm = gc->comm;
        const int rank = comm->rank;
        Grid* grids[2] = {ga, gc};
        const void* user[2] = {a, c};
        const int* desc[2] = {desca, descc};
        void* dev[2] = {const_cast<void*>(a), c};
        bool staged[2] = {false, false}, in_grid[2];
        size_t local_bytes[2] = {0, 0};
        const bool c_whole = ic == 1 && jc == 1 && m == descc[2] && n == descc[3];
        for (int x = 0; x < 2; ++x) {
            Grid* g = grids[x];
            in_grid[x] = rank < g->nprow * g->npcol;
            if (!in_grid[x] || !user[x] || !is_host_pointer(user[x])) continue;
            int myrow = 0, mycol = 0;
            costa::rank_to_grid(rank, g->nprow, g->npcol, g->order, &myrow, &mycol);
            const int loc_cols = costa::numroc(desc[x][3], desc[x][5], mycol, desc[x][7], g->npcol);
            local_bytes[x] = static_cast<size_t>(desc[x][8]) * std::max(loc_cols, 0) * eb;
            if (local_bytes[x] == 0) continue;
            const int slot = x == 0 ? 0 : 2;  // staging slots of the C grid: [0] source, [2] destination
            if (gc->stage_bytes[slot] < local_bytes[x]) {
                if (gc->stage[slot]) { cudaStreamSynchronize(stream); cudaFree(gc->stage[slot]); gc->stage[slot] = nullptr; gc->stage_bytes[slot] = 0; }
                if (cudaMalloc(reinterpret_cast<void**>(&gc->stage[slot]), local_bytes[x]) != cudaSuccess) {
                    set_last_error("p?tran / p?gemr2d: cudaMalloc of a staging buffer failed");
                    return COSMA_B200_OUT_OF_MEMORY;
                }
                gc->stage_bytes[slot] = local_bytes[x];
            }
            staged[x] = true;
            dev[x] = gc->stage[slot];
            const bool upload = x == 0 || !(is_zero(beta, cplx) && c_whole);
            if (upload) COSMA_B200_CUDA_TRY(cudaMemcpyAsync(dev[x], user[x], local_bytes[x], cudaMemcpyHostToDevice, stream));
        }
        const int a_subm = op == 'N' ? m : n, a_subn = op == 'N' ? n : m;
        costa::erased_layout LA = costa::erased_scalapack_layout(desca[8], desca[2], desca[3], ia, ja, a_subm, a_subn, desca[4], desca[5], ga->nprow,
                                                            ga->npcol, ga->order, desca[6], desca[7], dev[0], eb, 'C', in_grid[0] ? rank : -1);
        costa::erased_layout LC = costa::erased_scalapack_layout(descc[8], descc[2], descc[3], ic, jc, m, n, descc[4], descc[5], gc->nprow, gc->npcol,
                                                            gc->order, descc[6], descc[7], dev[1], eb, 'C', in_grid[1] ? rank : -1);
        LA.grid.n_ranks = LC.grid.n_ranks = comm->size;
        std::string key;
        key.push_back(dtype); key.push_back(op);
        append(key, alpha, (cplx ? 2 : 1) * sizeof(double));
        append(key, beta, (cplx ? 2 : 1) * sizeof(double));
        append_layout(key, LA);
        append_layout(key, LC);
        Grid::Cached* entry = nullptr;
        for (auto& e : gc->transforms)
            if (e.key == key) entry = &e;
        if (!entry) {
            if (gc->transforms.size() >= kMaxCachedTransforms) {
                size_t oldest = 0;
                for (size_t i = 1; i < gc->transforms.size(); ++i)
                    if (gc->transforms[i].stamp < gc->transforms[oldest].stamp) oldest = i;
                cudaStreamSynchronize(stream);  // the evicted plan may still be running
                gc->transforms.erase(gc->transforms.begin() + oldest);
            }
            Grid::Cached e;
            e.key = key;
            std::vector<costa::transform_spec> spec(1);
            spec[0].from = &LA; spec[0].to = &LC; spec[0].op = op;
            spec[0].alpha[0] = alpha[0]; spec[0].alpha[1] = cplx ? alpha[1] : 0.0;
            spec[0].beta[0] = beta[0]; spec[0].beta[1] = cplx ? beta[1] : 0.0;
            int rc = transform_plan_build(comm, rank, comm->size, dtype, spec, e.plan);
            if (rc != COSMA_B200_OK) return rc;
            gc->transforms.push_back(std::move(e));
            entry = &gc->transforms.back();
        }
        entry->stamp = ++gc->clock;
        int rc = transform_plan_run(*entry->plan, stream);
        if (rc != COSMA_B200_OK) {  // never keep a plan that failed (see layout_multiply)
            cudaStreamSynchronize(stream);
            for (size_t i = 0; i < gc->transforms.size(); ++i)
                if (&gc->transforms[i] == entry) {
                    gc->transforms.erase(gc->transforms.begin() + i);
                    break;
                }
            return rc;
        }
        gc->last_launches = entry->plan->last_launches;
        if (staged[1]) {
            int myrow = 0, mycol = 0;
            costa::rank_to_grid(rank, gc->nprow, gc->npcol, gc->order, &myrow, &mycol);
            COSMA_B200_CUDA_TRY(download_local(c, dev[1], descc[8], costa::numroc(descc[2], descc[4], myrow, descc[6], gc->nprow),
                                               costa::numroc(descc[3], descc[5], mycol, descc[7], gc->npcol), eb, stream));
        }
        return COSMA_B200_OK;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
}

int cosma_b200_pxtran(void* grid, char dtype, char op, int m, int n, const double* alpha, const void* a, int ia, int ja, const int* desca,
                      const double* beta, void* c, int ic, int jc, const int* descc, void* stream) {
    op = std::toupper(op);
    if (op != 'T' && op != 'C') return COSMA_B200_INVALID_ARG;
    return xptransform(static_cast<Grid*>(grid), static_cast<Grid*>(grid), dtype, op, m, n, alpha, a, ia, ja, desca, beta, c, ic, jc, descc,
                       static_cast<cudaStream_t>(stream));
}
int cosma_b200_pxgemr2d(void* grid_a, void* grid_c, char dtype, int m, int n, const void* a, int ia, int ja, const int* desca, void* c, int ic,
                        int jc, const int* descc, void* stream) {
    const double one[2] = {1.0, 0.0}, zero[2] = {0.0, 0.0};
    return xptransform(static_cast<Grid*>(grid_a), static_cast<Grid*>(grid_c), dtype, 'N', m, n, one, a, ia, ja, desca, zero, c, ic, jc, descc,
                       static_cast<cudaStream_t>(stream));
}
int cosma_b200_grid_last_launches(void* grid) { return grid ? static_cast<Grid*>(grid)->last_launches : 0; }

#define COSMA_B200_PTRAN(NAME, T, DT, OP)                                                                                              \
    int cosma_b200_##NAME(void* grid, int m, int n, const T* alpha, const T* a, int ia, int ja, const int* desca, const T* beta, T* c,   \
                          int ic, int jc, const int* descc, void* stream) {                                                            \
        if (!alpha || !beta) return COSMA_B200_INVALID_ARG;                                                                           \
        const bool cplx = DT == 'c' || DT == 'z';                                                                                     \
        const double a2[2] = {double(alpha[0]), cplx ? double(alpha[1]) : 0.0}, b2[2] = {double(beta[0]), cplx ? double(beta[1]) : 0.0}; \
        return cosma_b200_pxtran(grid, DT, OP, m, n, a2, a, ia, ja, desca, b2, c, ic, jc, descc, stream);                             \
    }
COSMA_B200_PTRAN(pstran, float, 's', 'T')
COSMA_B200_PTRAN(pdtran, double, 'd', 'T')
COSMA_B200_PTRAN(pctranu, float, 'c', 'T')
COSMA_B200_PTRAN(pztranu, double, 'z', 'T')
COSMA_B200_PTRAN(pctranc, float, 'c', 'C')
COSMA_B200_PTRAN(pztranc, double, 'z', 'C')
#undef COSMA_B200_PTRAN

#define COSMA_B200_PGEMR2D(NAME, T, DT)                                                                                               \
    int cosma_b200_##NAME(void* grid_a, void* grid_c, int m, int n, const T* a, int ia, int ja, const int* desca, T* c, int ic, int jc, \
                          const int* descc, void* stream) {                                                                            \
        return cosma_b200_pxgemr2d(grid_a, grid_c, DT, m, n, a, ia, ja, desca, c, ic, jc, descc, stream);                             \
    }
COSMA_B200_PGEMR2D(psgemr2d, float, 's')
COSMA_B200_PGEMR2D(pdgemr2d, double, 'd')
COSMA_B200_PGEMR2D(pcgemr2d, float, 'c')
COSMA_B200_PGEMR2D(pzgemr2d, double, 'z')
#undef COSMA_B200_PGEMR2D

}  // extern "C"
