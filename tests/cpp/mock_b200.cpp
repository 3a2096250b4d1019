// TEST INFRASTRUCTURE ONLY -- a CPU stand-in for the entry points of libcosma_b200.so that the C++ host layer calls,
// LD_PRELOADed in front of the real library by tests/test_z_cpp_api.py::test_cpp_programs_multirank_on_cpu. It lets the whole host
// layer (cosma::multiply / CosmaMatrix / multiply_using_layout, costa::transform, the C interface, cosma::pxgemm + BLACS-lite,
// the MPI-name subset, the test programs themselves) run on 1..16 RANKS on a box without GPUs: rank bookkeeping, idle ranks,
// strategies, coordinate maps, layout conversions and message protocols are all real; only the arithmetic is replaced by
// "gather the operands on the first rank, naive GEMM / dense relayout, scatter the result". It exists because a deadlock in
// the test protocol once burnt a round's GPU budget; it is never part of the product and never used on a GPU box.
#include <cosma/mapper.hpp>
#include <cosma/process_group.hpp>
#include <cosma/strategy.hpp>
#include <cosma_b200.h>
#include <costa/erased_layout.hpp>

#include <algorithm>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace pg = cosma::pg;

namespace {

thread_local std::string g_err;
int fail(const std::exception& e) {
    g_err = e.what();
    return COSMA_B200_INVALID_ARG;
}

int elem_bytes(char dtype) { return dtype == 's' ? 4 : dtype == 'd' ? 8 : dtype == 'c' ? 8 : dtype == 'z' ? 16 : 0; }

// ---- communicator: the members' world ranks, discovered through the leader whose world rank travels in the "unique id" ----
struct MockComm {
    int rank = 0, size = 1;
    std::vector<int> world;  // world rank of every member
    int tag = 0;             // base tag of this communicator's traffic on the world group
    void send(const void* buf, size_t bytes, int dst, int sub) const { pg::send(pg::world(), buf, bytes, world[dst], tag - sub); }
    void recv(void* buf, size_t bytes, int src, int sub) const { pg::recv(pg::world(), buf, bytes, world[src], tag - sub); }
};

struct mock_id {
    std::uint32_t magic, leader_world, counter, pad;
};

int base_tag(const mock_id& id) { return -(100000 + static_cast<int>((id.leader_world * 1009u + id.counter) % 20000u) * 16); }

// ---- dense <-> distributed through a layout (all on the leader) -----------------------------------------------------------
struct LayoutCopy {  // deep copy of a cosma_b200_layout + ordering
    costa::erased_layout l;
};

LayoutCopy copy_layout(const cosma_b200_layout& c, char ordering, int nranks) {
    LayoutCopy out;
    out.l.ordering = ordering;
    out.l.grid.grid.rows_split.assign(c.rowsplit, c.rowsplit + c.rowblocks + 1);
    out.l.grid.grid.cols_split.assign(c.colsplit, c.colsplit + c.colblocks + 1);
    out.l.grid.owners.assign(c.owners, c.owners + static_cast<size_t>(c.rowblocks) * c.colblocks);
    out.l.grid.n_ranks = nranks;
    for (int b = 0; b < c.nlocalblocks; ++b)
        out.l.blocks.push_back(costa::local_block{c.localblocks[b].row, c.localblocks[b].col, c.localblocks[b].data, c.localblocks[b].ld});
    return out;
}

const costa::local_block* find_block(const costa::erased_layout& l, int bi, int bj) {
    for (const auto& b : l.blocks)
        if (b.bi == bi && b.bj == bj) return &b;
    return nullptr;
}

template <typename T>
T& at(const costa::erased_layout& l, const costa::local_block& b, int li, int lj) {
    T* p = static_cast<T*>(b.data);
    return l.ordering == 'R' ? p[static_cast<size_t>(li) * b.ld + lj] : p[static_cast<size_t>(lj) * b.ld + li];
}

// leader ends up with the column-major dense matrix; everybody else sends its blocks in grid order
template <typename T>
std::vector<T> gather_dense(const costa::erased_layout& l, const MockComm& c, int sub) {
    const auto& g = l.grid.grid;
    const int rows = g.total_rows();
    std::vector<T> dense;
    if (c.rank == 0) dense.assign(static_cast<size_t>(rows) * g.total_cols(), T{0});
    for (int bi = 0; bi < g.n_rows(); ++bi)
        for (int bj = 0; bj < g.n_cols(); ++bj) {
            const int owner = l.grid.owner(bi, bj), r0 = g.rows_split[bi], c0 = g.cols_split[bj];
            const int nr = g.rows_split[bi + 1] - r0, nc = g.cols_split[bj + 1] - c0;
            if (nr <= 0 || nc <= 0 || owner < 0 || owner >= c.size) continue;
            if (owner != c.rank && c.rank != 0) continue;
            std::vector<T> packed(static_cast<size_t>(nr) * nc);
            if (owner == c.rank) {
                const costa::local_block* b = find_block(l, bi, bj);
                if (!b) throw std::runtime_error("mock: a block owned by this rank is missing from its local block list");
                for (int j = 0; j < nc; ++j)
                    for (int i = 0; i < nr; ++i) packed[static_cast<size_t>(j) * nr + i] = at<T>(l, *b, i, j);
                if (c.rank != 0) c.send(packed.data(), packed.size() * sizeof(T), 0, sub);
            } else {
                c.recv(packed.data(), packed.size() * sizeof(T), owner, sub);
            }
            if (c.rank == 0)
                for (int j = 0; j < nc; ++j)
                    for (int i = 0; i < nr; ++i) dense[static_cast<size_t>(c0 + j) * rows + r0 + i] = packed[static_cast<size_t>(j) * nr + i];
        }
    return dense;
}

template <typename T>
void scatter_dense(const costa::erased_layout& l, const MockComm& c, const std::vector<T>& dense, int sub) {
    const auto& g = l.grid.grid;
    const int rows = g.total_rows();
    for (int bi = 0; bi < g.n_rows(); ++bi)
        for (int bj = 0; bj < g.n_cols(); ++bj) {
            const int owner = l.grid.owner(bi, bj), r0 = g.rows_split[bi], c0 = g.cols_split[bj];
            const int nr = g.rows_split[bi + 1] - r0, nc = g.cols_split[bj + 1] - c0;
            if (nr <= 0 || nc <= 0 || owner < 0 || owner >= c.size) continue;
            if (owner != c.rank && c.rank != 0) continue;
            std::vector<T> packed(static_cast<size_t>(nr) * nc);
            if (c.rank == 0) {
                for (int j = 0; j < nc; ++j)
                    for (int i = 0; i < nr; ++i) packed[static_cast<size_t>(j) * nr + i] = dense[static_cast<size_t>(c0 + j) * rows + r0 + i];
                if (owner != 0) { c.send(packed.data(), packed.size() * sizeof(T), owner, sub); continue; }
            } else {
                c.recv(packed.data(), packed.size() * sizeof(T), 0, sub);
            }
            const costa::local_block* b = find_block(l, bi, bj);
            if (!b) throw std::runtime_error("mock: a block owned by this rank is missing from its local block list");
            for (int j = 0; j < nc; ++j)
                for (int i = 0; i < nr; ++i) at<T>(l, *b, i, j) = packed[static_cast<size_t>(j) * nr + i];
        }
}

template <typename T> T conj_of(const T& v) { return v; }
template <typename T> std::complex<T> conj_of(const std::complex<T>& v) { return std::conj(v); }

// op(X) of a column-major rows x cols matrix
template <typename T>
std::vector<T> apply_op(const std::vector<T>& x, int rows, int cols, char op) {
    if (op == 'N') return x;
    std::vector<T> out(x.size());
    for (int j = 0; j < cols; ++j)
        for (int i = 0; i < rows; ++i) out[static_cast<size_t>(i) * cols + j] = op == 'C' ? conj_of(x[static_cast<size_t>(j) * rows + i]) : x[static_cast<size_t>(j) * rows + i];
    return out;
}

// naive C = alpha * A * B + beta * C in the summation order of the reference's local_multiply_cpu; beta == 0 does not read C
template <typename T>
void naive_gemm(int m, int n, int k, T alpha, const T* A, const T* B, T beta, T* C) {
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) {
            T acc = beta == T{0} ? T{0} : C[static_cast<size_t>(j) * m + i] * beta;
            for (int p = 0; p < k; ++p) acc += alpha * A[static_cast<size_t>(p) * m + i] * B[static_cast<size_t>(j) * k + p];
            C[static_cast<size_t>(j) * m + i] = acc;
        }
}

template <typename T> T scalar_of(const double* v) { return static_cast<T>(v[0]); }
template <> std::complex<double> scalar_of<std::complex<double>>(const double* v) { return {v[0], v[1]}; }
template <> std::complex<float> scalar_of<std::complex<float>>(const double* v) { return {static_cast<float>(v[0]), static_cast<float>(v[1])}; }

template <typename T>
void layout_multiply_t(const MockComm& c, char ta, char tb, const double* alpha, const double* beta, const costa::erased_layout& A,
                       const costa::erased_layout& B, const costa::erased_layout& C) {
    const int m = C.num_rows(), n = C.num_cols(), k = ta == 'N' ? A.num_cols() : A.num_rows();
    if (m == 0 || n == 0) return;
    const T al = scalar_of<T>(alpha), be = scalar_of<T>(beta);
    std::vector<T> dA, dB, dC;
    const bool product = k > 0 && al != T{0};
    if (product) {
        dA = gather_dense<T>(A, c, 1);
        dB = gather_dense<T>(B, c, 2);
    }
    if (be != T{0}) dC = gather_dense<T>(C, c, 3);
    else if (c.rank == 0) dC.assign(static_cast<size_t>(m) * n, T{0});
    if (c.rank == 0) {
        if (product) {
            const std::vector<T> oa = apply_op(dA, A.num_rows(), A.num_cols(), ta), ob = apply_op(dB, B.num_rows(), B.num_cols(), tb);
            naive_gemm<T>(m, n, k, al, oa.data(), ob.data(), be, dC.data());
        } else {
            for (auto& v : dC) v = be == T{0} ? T{0} : be * v;
        }
    }
    scatter_dense<T>(C, c, dC, 4);
}

void layout_multiply(char dtype, const MockComm& c, char ta, char tb, const double* alpha, const double* beta, const costa::erased_layout& A,
                     const costa::erased_layout& B, const costa::erased_layout& C) {
    switch (dtype) {
        case 's': layout_multiply_t<float>(c, ta, tb, alpha, beta, A, B, C); break;
        case 'd': layout_multiply_t<double>(c, ta, tb, alpha, beta, A, B, C); break;
        case 'c': layout_multiply_t<std::complex<float>>(c, ta, tb, alpha, beta, A, B, C); break;
        default: layout_multiply_t<std::complex<double>>(c, ta, tb, alpha, beta, A, B, C); break;
    }
}

template <typename T>
void transform_t(const MockComm& c, const costa::erased_layout& F, const costa::erased_layout& G, char op, const double* alpha, const double* beta) {
    const T al = scalar_of<T>(alpha), be = scalar_of<T>(beta);
    std::vector<T> src = gather_dense<T>(F, c, 5), dst;
    if (be != T{0}) dst = gather_dense<T>(G, c, 6);
    if (c.rank == 0) {
        const std::vector<T> o = apply_op(src, F.num_rows(), F.num_cols(), op);
        if (be == T{0}) dst.assign(o.size(), T{0});
        for (size_t i = 0; i < o.size(); ++i) {
            if (al == T{1} && be == T{0}) dst[i] = o[i];
            else dst[i] = be == T{0} ? al * o[i] : be * dst[i] + al * o[i];
        }
    }
    scatter_dense<T>(G, c, dst, 7);
}

struct MockPlan {
    MockComm* comm = nullptr;
    cosma::Strategy strategy;
    char dtype = 'd';
    int rank = 0;
};

template <typename T>
void multiply_native_t(MockPlan& p, const double* alpha, const double* beta, const void* A_, const void* B_, void* C_) {
    const cosma::Strategy& s = p.strategy;
    const int P = static_cast<int>(s.P);
    if (p.rank >= P) return;
    MockComm& c = *p.comm;
    const T al = scalar_of<T>(alpha), be = scalar_of<T>(beta);
    const void* local[3] = {A_, B_, C_};
    const int dims[3][2] = {{s.m, s.k}, {s.k, s.n}, {s.m, s.n}};
    std::vector<T> dense[3];
    std::vector<std::unique_ptr<cosma::Mapper>> mappers;
    for (int x = 0; x < 3; ++x) mappers.emplace_back(new cosma::Mapper("ABC"[x], s, p.rank));
    auto place = [&](int x, int r, std::vector<T>& buf, bool to_dense) {
        const auto& blocks = mappers[x]->initial_layout(r);
        const auto& offs = mappers[x]->blocks_offsets(r);
        for (size_t b = 0; b < blocks.size(); ++b) {
            const int r0 = blocks[b].rows.first(), c0 = blocks[b].cols.first();
            const int nr = static_cast<int>(blocks[b].rows.length()), nc = static_cast<int>(blocks[b].cols.length());
            for (int j = 0; j < nc; ++j)
                for (int i = 0; i < nr; ++i) {
                    T& d = dense[x][static_cast<size_t>(c0 + j) * dims[x][0] + r0 + i];
                    T& l = buf[offs[b] + static_cast<size_t>(j) * nr + i];
                    if (to_dense) d = l; else l = d;
                }
        }
    };
    for (int x = 0; x < 3; ++x) {
        if (x == 2 && be == T{0}) {
            if (c.rank == 0) dense[2].assign(static_cast<size_t>(s.m) * s.n, T{0});
            continue;
        }
        const size_t mine = mappers[x]->initial_size(p.rank);
        if (c.rank == 0) {
            dense[x].assign(static_cast<size_t>(dims[x][0]) * dims[x][1], T{0});
            for (int r = 0; r < P; ++r) {
                std::vector<T> buf(mappers[x]->initial_size(r));
                if (r == 0) std::memcpy(buf.data(), local[x], buf.size() * sizeof(T));
                else c.recv(buf.data(), buf.size() * sizeof(T), r, 8 + x);
                place(x, r, buf, true);
            }
        } else {
            c.send(local[x], mine * sizeof(T), 0, 8 + x);
        }
    }
    if (c.rank == 0) naive_gemm<T>(s.m, s.n, s.k, al, dense[0].data(), dense[1].data(), be, dense[2].data());
    if (c.rank == 0) {
        for (int r = 0; r < P; ++r) {
            std::vector<T> buf(mappers[2]->initial_size(r));
            place(2, r, buf, false);
            if (r == 0) std::memcpy(C_, buf.data(), buf.size() * sizeof(T));
            else c.send(buf.data(), buf.size() * sizeof(T), r, 11);
        }
    } else {
        c.recv(C_, mappers[2]->initial_size(p.rank) * sizeof(T), 0, 11);
    }
}

struct MockTransform {
    MockComm* comm = nullptr;
    char dtype = 'd';
    std::vector<LayoutCopy> from, to;
    std::string ops;
    std::vector<double> alpha, beta;
};

struct MockGrid {
    MockComm* comm = nullptr;
    char order = 'R';
    int nprow = 1, npcol = 1;
};

std::uint32_t g_counter = 0;

// the stand-in moves its data over cosma::pg: bring the group up when the library is loaded, on every rank at once (with the
// MPI-name subset MPI_Init would do it; a build against a real MPI -- tests run that over the reference checker's minimpi -- would
// otherwise first touch pg on whichever rank calls the C ABI first, and wait for the others forever)
struct pg_bootstrap {
    pg_bootstrap() {
        try { pg::init(); } catch (const std::exception&) {}
    }
} g_pg_bootstrap;

}  // namespace

extern "C" {

const char* cosma_b200_version(void) { return "cosma_b200 MOCK (tests/cpp/mock_b200.cpp): CPU stand-in, test infrastructure"; }
const char* cosma_b200_last_error(void) { return g_err.c_str(); }

int cosma_b200_device_count(int* count) { *count = 1; return COSMA_B200_OK; }
int cosma_b200_set_device(int) { return COSMA_B200_OK; }
int cosma_b200_device_pci_bus_id(int, char* out, int out_len) {  // MOCK_PCI_BDF: a PCI function that exists under /sys/bus/pci/devices
    const char* v = std::getenv("MOCK_PCI_BDF");
    if (!v || !out || static_cast<int>(std::strlen(v)) + 1 > out_len) return COSMA_B200_CUDA_ERROR;
    std::strcpy(out, v);
    return COSMA_B200_OK;
}
int cosma_b200_stream_synchronize(void*) { return COSMA_B200_OK; }
int cosma_b200_host_alloc(void** ptr, uint64_t bytes) {
    *ptr = bytes ? std::malloc(bytes) : nullptr;
    return (*ptr || !bytes) ? COSMA_B200_OK : COSMA_B200_OUT_OF_MEMORY;
}
int cosma_b200_host_free(void* ptr) { std::free(ptr); return COSMA_B200_OK; }
int cosma_b200_host_register(void*, uint64_t) { return COSMA_B200_OK; }
int cosma_b200_host_unregister(void*) { return COSMA_B200_OK; }

int cosma_b200_nccl_unique_id(uint8_t* out128) {
    std::memset(out128, 0, 128);
    mock_id id{0xC05A5A5Au, static_cast<std::uint32_t>(pg::rank(pg::world())), ++g_counter, 0};
    std::memcpy(out128, &id, sizeof(id));
    return COSMA_B200_OK;
}

int cosma_b200_comm_create(int rank, int nranks, const uint8_t* id128, void** comm_out) {
    try {
        auto c = std::make_unique<MockComm>();
        c->rank = rank;
        c->size = nranks;
        const int me = pg::rank(pg::world());
        if (nranks == 1) {
            c->world = {me};
        } else {
            mock_id id;
            std::memcpy(&id, id128, sizeof(id));
            if (id.magic != 0xC05A5A5Au) throw std::runtime_error("mock: unique id did not come from the mock");
            c->tag = base_tag(id);
            c->world.assign(nranks, -1);
            std::int32_t hello[2] = {me, rank};
            if (rank == 0) {
                c->world[0] = me;
                for (int i = 1; i < nranks; ++i) {
                    std::int32_t h[2];
                    pg::recv_any(pg::world(), h, sizeof(h), c->tag - 15);
                    c->world[h[1]] = h[0];
                }
                for (int i = 1; i < nranks; ++i) pg::send(pg::world(), c->world.data(), sizeof(int) * nranks, c->world[i], c->tag - 14);
            } else {
                pg::send(pg::world(), hello, sizeof(hello), static_cast<int>(id.leader_world), c->tag - 15);
                pg::recv(pg::world(), c->world.data(), sizeof(int) * nranks, static_cast<int>(id.leader_world), c->tag - 14);
            }
        }
        *comm_out = c.release();
        return COSMA_B200_OK;
    } catch (const std::exception& e) { return fail(e); }
}
int cosma_b200_comm_destroy(void* comm) { delete static_cast<MockComm*>(comm); return COSMA_B200_OK; }

int cosma_b200_plan_create_for_strategy(void* comm, int rank, int nranks, int m, int n, int k, int P, const char* steps, char dtype, void** plan_out) {
    try {
        auto p = std::make_unique<MockPlan>();
        p->comm = static_cast<MockComm*>(comm);
        p->dtype = dtype;
        p->rank = p->comm ? p->comm->rank : rank;
        const std::string st = steps ? steps : "";
        if (st.find_first_not_of(" ,") == std::string::npos) {
            std::vector<int> divs;
            std::string dims, types;
            p->strategy = cosma::Strategy(m, n, k, static_cast<size_t>(P), divs, dims, types);
        } else {
            p->strategy = cosma::parse_strategy(m, n, k, static_cast<size_t>(P), st);
        }
        if (P > (p->comm ? p->comm->size : nranks)) throw std::runtime_error("mock: the strategy uses more ranks than the communicator has");
        *plan_out = p.release();
        return COSMA_B200_OK;
    } catch (const std::exception& e) { return fail(e); }
}
int cosma_b200_plan_destroy(void* plan) { delete static_cast<MockPlan*>(plan); return COSMA_B200_OK; }
int64_t cosma_b200_plan_arena_elements(void* plan, int matrix) {
    MockPlan* p = static_cast<MockPlan*>(plan);
    if (p->rank >= static_cast<int>(p->strategy.P)) return 0;
    return static_cast<int64_t>(cosma::Mapper("ABC"[matrix], p->strategy, p->rank).initial_size(p->rank));
}

int cosma_b200_multiply_host(void* plan, const double* alpha, const double* beta, const void* A, const void* B, void* C, void*) {
    try {
        MockPlan& p = *static_cast<MockPlan*>(plan);
        switch (p.dtype) {
            case 's': multiply_native_t<float>(p, alpha, beta, A, B, C); break;
            case 'd': multiply_native_t<double>(p, alpha, beta, A, B, C); break;
            case 'c': multiply_native_t<std::complex<float>>(p, alpha, beta, A, B, C); break;
            default: multiply_native_t<std::complex<double>>(p, alpha, beta, A, B, C); break;
        }
        return COSMA_B200_OK;
    } catch (const std::exception& e) { return fail(e); }
}

static int mock_layout_multiply(void* comm, char dtype, const char* transa, const char* transb, const double* alpha, const cosma_b200_layout* A,
                                const cosma_b200_layout* B, const double* beta, const cosma_b200_layout* C) {
    try {
        MockComm& c = *static_cast<MockComm*>(comm);
        const LayoutCopy a = copy_layout(*A, 'C', c.size), b = copy_layout(*B, 'C', c.size), cc = copy_layout(*C, 'C', c.size);
        const bool cplx = dtype == 'c' || dtype == 'z';
        const double a2[2] = {alpha[0], cplx ? alpha[1] : 0.0}, b2[2] = {beta[0], cplx ? beta[1] : 0.0};
        layout_multiply(dtype, c, static_cast<char>(std::toupper(*transa)), static_cast<char>(std::toupper(*transb)), a2, b2, a.l, b.l, cc.l);
        return COSMA_B200_OK;
    } catch (const std::exception& e) { return fail(e); }
}
int cosma_b200_smultiply_using_layout(void* comm, const char* ta, const char* tb, const double* alpha, const cosma_b200_layout* A,
                                      const cosma_b200_layout* B, const double* beta, const cosma_b200_layout* C, void*) {
    return mock_layout_multiply(comm, 's', ta, tb, alpha, A, B, beta, C);
}
int cosma_b200_dmultiply_using_layout(void* comm, const char* ta, const char* tb, const double* alpha, const cosma_b200_layout* A,
                                      const cosma_b200_layout* B, const double* beta, const cosma_b200_layout* C, void*) {
    return mock_layout_multiply(comm, 'd', ta, tb, alpha, A, B, beta, C);
}
int cosma_b200_cmultiply_using_layout(void* comm, const char* ta, const char* tb, const double* alpha, const cosma_b200_layout* A,
                                      const cosma_b200_layout* B, const double* beta, const cosma_b200_layout* C, void*) {
    return mock_layout_multiply(comm, 'c', ta, tb, alpha, A, B, beta, C);
}
int cosma_b200_zmultiply_using_layout(void* comm, const char* ta, const char* tb, const double* alpha, const cosma_b200_layout* A,
                                      const cosma_b200_layout* B, const double* beta, const cosma_b200_layout* C, void*) {
    return mock_layout_multiply(comm, 'z', ta, tb, alpha, A, B, beta, C);
}

int cosma_b200_transform_plan_create(void* comm, int, int, char dtype, int n, const cosma_b200_layout* from, const cosma_b200_layout* to,
                                     const char* ordering_from, const char* ordering_to, const char* trans, const double* alpha, const double* beta,
                                     void** plan_out) {
    try {
        auto t = std::make_unique<MockTransform>();
        t->comm = static_cast<MockComm*>(comm);
        t->dtype = dtype;
        for (int i = 0; i < n; ++i) {
            t->from.push_back(copy_layout(from[i], ordering_from ? ordering_from[i] : 'C', t->comm->size));
            t->to.push_back(copy_layout(to[i], ordering_to ? ordering_to[i] : 'C', t->comm->size));
            t->ops.push_back(trans ? static_cast<char>(std::toupper(trans[i])) : 'N');
            t->alpha.push_back(alpha ? alpha[2 * i] : 1.0); t->alpha.push_back(alpha ? alpha[2 * i + 1] : 0.0);
            t->beta.push_back(beta ? beta[2 * i] : 0.0); t->beta.push_back(beta ? beta[2 * i + 1] : 0.0);
            if (dtype == 's' || dtype == 'd') t->alpha[2 * i + 1] = t->beta[2 * i + 1] = 0.0;
        }
        *plan_out = t.release();
        return COSMA_B200_OK;
    } catch (const std::exception& e) { return fail(e); }
}
int cosma_b200_transform_run(void* plan, void*) {
    try {
        MockTransform& t = *static_cast<MockTransform*>(plan);
        for (size_t i = 0; i < t.from.size(); ++i) {
            const double* a = &t.alpha[2 * i];
            const double* b = &t.beta[2 * i];
            switch (t.dtype) {
                case 's': transform_t<float>(*t.comm, t.from[i].l, t.to[i].l, t.ops[i], a, b); break;
                case 'd': transform_t<double>(*t.comm, t.from[i].l, t.to[i].l, t.ops[i], a, b); break;
                case 'c': transform_t<std::complex<float>>(*t.comm, t.from[i].l, t.to[i].l, t.ops[i], a, b); break;
                default: transform_t<std::complex<double>>(*t.comm, t.from[i].l, t.to[i].l, t.ops[i], a, b); break;
            }
        }
        return COSMA_B200_OK;
    } catch (const std::exception& e) { return fail(e); }
}
int cosma_b200_transform_plan_destroy(void* plan) { delete static_cast<MockTransform*>(plan); return COSMA_B200_OK; }

int cosma_b200_grid_create(void* comm, char order, int nprow, int npcol, void** grid_out) {
    auto g = new MockGrid;
    g->comm = static_cast<MockComm*>(comm);
    g->order = static_cast<char>(std::toupper(order));
    g->nprow = nprow;
    g->npcol = npcol;
    *grid_out = g;
    return COSMA_B200_OK;
}
int cosma_b200_grid_destroy(void* grid) { delete static_cast<MockGrid*>(grid); return COSMA_B200_OK; }

static int mock_pgemm(void* grid, char dtype, char ta, char tb, int m, int n, int k, const double* alpha, const void* a, int ia, int ja, const int* da,
                      const void* b, int ib, int jb, const int* db, const double* beta, void* c, int ic, int jc, const int* dc) {
    try {
        MockGrid& g = *static_cast<MockGrid*>(grid);
        if (m == 0 || n == 0) return COSMA_B200_OK;
        ta = static_cast<char>(std::toupper(ta));
        tb = static_cast<char>(std::toupper(tb));
        const int eb = elem_bytes(dtype), rank = g.comm->rank;
        const bool in_grid = rank < g.nprow * g.npcol;
        auto layout_of = [&](const int* d, const void* ptr, int i0, int j0, int sm, int sn) {
            costa::erased_layout l = costa::erased_scalapack_layout(d[8], d[2], d[3], i0, j0, sm, sn, d[4], d[5], g.nprow, g.npcol, g.order, d[6], d[7],
                                                                    const_cast<void*>(ptr), eb, 'C', in_grid ? rank : -1);
            l.grid.n_ranks = g.comm->size;
            return l;
        };
        const bool cplx = dtype == 'c' || dtype == 'z';
        const double a2[2] = {alpha[0], cplx ? alpha[1] : 0.0}, b2[2] = {beta[0], cplx ? beta[1] : 0.0};
        // k == 0: op(A) is m x 0; give the multiply an empty product by zeroing alpha
        const bool scale_only = k == 0 || (a2[0] == 0.0 && a2[1] == 0.0);
        const costa::erased_layout LC = layout_of(dc, c, ic, jc, m, n);
        if (scale_only) {
            const double zero[2] = {0.0, 0.0};
            costa::erased_layout dummyA = LC, dummyB = LC;  // never gathered when alpha == 0
            layout_multiply(dtype, *g.comm, 'N', 'N', zero, b2, dummyA, dummyB, LC);
        } else {
            const costa::erased_layout LA = layout_of(da, a, ia, ja, ta == 'N' ? m : k, ta == 'N' ? k : m);
            const costa::erased_layout LB = layout_of(db, b, ib, jb, tb == 'N' ? k : n, tb == 'N' ? n : k);
            layout_multiply(dtype, *g.comm, ta, tb, a2, b2, LA, LB, LC);
        }
        return COSMA_B200_OK;
    } catch (const std::exception& e) { return fail(e); }
}
int cosma_b200_psgemm(void* grid, char ta, char tb, int m, int n, int k, const double* alpha, const float* a, int ia, int ja, const int* da, const float* b,
                      int ib, int jb, const int* db, const double* beta, float* c, int ic, int jc, const int* dc, void*) {
    return mock_pgemm(grid, 's', ta, tb, m, n, k, alpha, a, ia, ja, da, b, ib, jb, db, beta, c, ic, jc, dc);
}
int cosma_b200_pdgemm(void* grid, char ta, char tb, int m, int n, int k, const double* alpha, const double* a, int ia, int ja, const int* da,
                      const double* b, int ib, int jb, const int* db, const double* beta, double* c, int ic, int jc, const int* dc, void*) {
    return mock_pgemm(grid, 'd', ta, tb, m, n, k, alpha, a, ia, ja, da, b, ib, jb, db, beta, c, ic, jc, dc);
}
int cosma_b200_pcgemm(void* grid, char ta, char tb, int m, int n, int k, const double* alpha, const float* a, int ia, int ja, const int* da, const float* b,
                      int ib, int jb, const int* db, const double* beta, float* c, int ic, int jc, const int* dc, void*) {
    return mock_pgemm(grid, 'c', ta, tb, m, n, k, alpha, a, ia, ja, da, b, ib, jb, db, beta, c, ic, jc, dc);
}
int cosma_b200_pzgemm(void* grid, char ta, char tb, int m, int n, int k, const double* alpha, const double* a, int ia, int ja, const int* da,
                      const double* b, int ib, int jb, const int* db, const double* beta, double* c, int ic, int jc, const int* dc, void*) {
    return mock_pgemm(grid, 'z', ta, tb, m, n, k, alpha, a, ia, ja, da, b, ib, jb, db, beta, c, ic, jc, dc);
}

// p?tran / p?tranu / p?tranc and p?gemr2d: one dense relayout between two block-cyclic sub-matrices
int cosma_b200_pxtran(void* grid, char dtype, char op, int m, int n, const double* alpha, const void* a, int ia, int ja, const int* da, const double* beta,
                      void* c, int ic, int jc, const int* dc, void*) {
    try {
        MockGrid& g = *static_cast<MockGrid*>(grid);
        if (m == 0 || n == 0) return COSMA_B200_OK;
        const int eb = elem_bytes(dtype), rank = g.comm->rank;
        const bool in_grid = rank < g.nprow * g.npcol;
        costa::erased_layout LA = costa::erased_scalapack_layout(da[8], da[2], da[3], ia, ja, n, m, da[4], da[5], g.nprow, g.npcol, g.order, da[6], da[7],
                                                                 const_cast<void*>(a), eb, 'C', in_grid ? rank : -1);
        costa::erased_layout LC = costa::erased_scalapack_layout(dc[8], dc[2], dc[3], ic, jc, m, n, dc[4], dc[5], g.nprow, g.npcol, g.order, dc[6], dc[7], c,
                                                                 eb, 'C', in_grid ? rank : -1);
        LA.grid.n_ranks = LC.grid.n_ranks = g.comm->size;
        const bool cplx = dtype == 'c' || dtype == 'z';
        const double a2[2] = {alpha[0], cplx ? alpha[1] : 0.0}, b2[2] = {beta[0], cplx ? beta[1] : 0.0};
        op = static_cast<char>(std::toupper(op));
        switch (dtype) {
            case 's': transform_t<float>(*g.comm, LA, LC, op, a2, b2); break;
            case 'd': transform_t<double>(*g.comm, LA, LC, op, a2, b2); break;
            case 'c': transform_t<std::complex<float>>(*g.comm, LA, LC, op, a2, b2); break;
            default: transform_t<std::complex<double>>(*g.comm, LA, LC, op, a2, b2); break;
        }
        return COSMA_B200_OK;
    } catch (const std::exception& e) { return fail(e); }
}
int cosma_b200_pxgemr2d(void* grid_a, void* grid_c, char dtype, int m, int n, const void* a, int ia, int ja, const int* da, void* c, int ic, int jc,
                        const int* dc, void*) {
    try {
        MockGrid& ga = *static_cast<MockGrid*>(grid_a);
        MockGrid& gc = *static_cast<MockGrid*>(grid_c);
        if (m == 0 || n == 0) return COSMA_B200_OK;
        if (ga.comm != gc.comm) throw std::runtime_error("mock: p?gemr2d grids on different communicators");
        const int eb = elem_bytes(dtype), rank = ga.comm->rank;
        costa::erased_layout LA = costa::erased_scalapack_layout(da[8], da[2], da[3], ia, ja, m, n, da[4], da[5], ga.nprow, ga.npcol, ga.order, da[6], da[7],
                                                                 const_cast<void*>(a), eb, 'C', rank < ga.nprow * ga.npcol ? rank : -1);
        costa::erased_layout LC = costa::erased_scalapack_layout(dc[8], dc[2], dc[3], ic, jc, m, n, dc[4], dc[5], gc.nprow, gc.npcol, gc.order, dc[6], dc[7], c,
                                                                 eb, 'C', rank < gc.nprow * gc.npcol ? rank : -1);
        LA.grid.n_ranks = LC.grid.n_ranks = ga.comm->size;
        const double one[2] = {1.0, 0.0}, zero[2] = {0.0, 0.0};
        switch (dtype) {
            case 's': transform_t<float>(*ga.comm, LA, LC, 'N', one, zero); break;
            case 'd': transform_t<double>(*ga.comm, LA, LC, 'N', one, zero); break;
            case 'c': transform_t<std::complex<float>>(*ga.comm, LA, LC, 'N', one, zero); break;
            default: transform_t<std::complex<double>>(*ga.comm, LA, LC, 'N', one, zero); break;
        }
        return COSMA_B200_OK;
    } catch (const std::exception& e) { return fail(e); }
}

}  // extern "C"
