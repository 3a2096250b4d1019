// R3/R4: batched block copy / transpose (+conjugate, +alpha/beta) for sm_100a -- the device form of COSTA's
// copy_and_transform (reference libs/COSTA/src/costa/grid2grid/memory_utils.hpp:287-346: copy2D :47-85,
// transpose_col_major :88-166, transpose_row_major :169-250), which the reference runs on the host with OpenMP for
// every message of a layout transformation (communication_data.cpp:166-302).
//
// One launch handles a whole list of pieces. HBM-bound: 2 * elements * sizeof(T) algorithmic bytes per launch
// (3 * when beta != 0). Design:
//  * every piece is normalised on the host to a column-major view (row-major storage = the transposed column-major
//    matrix), so the four ordering combinations of the reference collapse into "copy" or "transpose";
//  * pieces are cut into 16 KB tiles (copies: one contiguous kilobyte per column x 16 columns; transposes: near-square
//    in bytes, 32x32 / 64x32 / 64x64 elements for 16 / 8 / 4-byte types), numbered globally through
//    a prefix sum kept in the descriptors; each CTA of a grid of exactly SMs x resident-CTAs takes a CONTIGUOUS range
//    of tiles (one binary search, then a linear walk), column-major inside a piece so consecutive tiles touch
//    neighbouring DRAM pages;
//  * every thread keeps four 16-byte requests in flight before its first store; the piece descriptor and scalars live in
//    shared memory, not registers, so six CTAs (1536 threads, 96 KB of requests) stay resident per SM;
//  * copies run along the contiguous dimension, 512 contiguous bytes per warp request; transposes use NO shared
//    memory: a thread owns VEC x VEC micro-tiles (VEC = 16 B / sizeof(T)), reads them as 16-byte requests along the
//    source rows, transposes in registers and writes 16-byte requests along the destination rows; lanes form an 8 x 4
//    grid so a warp reads 128 and writes 64 contiguous bytes per column -- whole sectors on both sides;
//  * pieces whose addresses or leading dimensions are not 16-byte aligned take the same code with element-sized
//    requests;
//  * alpha = 1, beta = 0 moves bits only (conjugation flips one sign bit) -> bit-exact; otherwise
//    dst = beta*dst + alpha*op(src) is evaluated with unfused multiplies and adds in the reference's order, and
//    beta == 0 never reads dst.
#include "relayout_sm100.h"
#include "../../include/cosma_b200.h"

#include <algorithm>
#include <cstdlib>
#include <string>

namespace cosma_b200 {
void set_last_error(const std::string& msg);

namespace {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int CHUNK = 4;                         // 16-byte requests a thread issues before its first store

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

// real element
template <typename R>
struct RealOps {
    using E = R;
    static __device__ __forceinline__ E conj(E v) { return v; }
    static __device__ __forceinline__ E zero() { return R(0); }
    static __device__ __forceinline__ E axpby(const DevScalars& s, E x, E d, bool read_dst) {
        const R ax = mul_rn(static_cast<R>(s.alpha[0]), x);
        return read_dst ? add_rn(mul_rn(static_cast<R>(s.beta[0]), d), ax) : ax;
    }
};
template <typename R, typename R2>
struct CplxOps {
    using E = R2;
    static __device__ __forceinline__ E conj(E v) { v.y = -v.y; return v; }
    static __device__ __forceinline__ E zero() { E z; z.x = R(0); z.y = R(0); return z; }
    static __device__ __forceinline__ E mul(R ar, R ai, E x) {
        E o;
        o.x = sub_rn(mul_rn(ar, x.x), mul_rn(ai, x.y));
        o.y = add_rn(mul_rn(ar, x.y), mul_rn(ai, x.x));
        return o;
    }
    static __device__ __forceinline__ E axpby(const DevScalars& s, E x, E d, bool read_dst) {
        const E ax = mul(static_cast<R>(s.alpha[0]), static_cast<R>(s.alpha[1]), x);
        if (!read_dst) return ax;
        const E bd = mul(static_cast<R>(s.beta[0]), static_cast<R>(s.beta[1]), d);
        E o;
        o.x = add_rn(bd.x, ax.x);
        o.y = add_rn(bd.y, ax.y);
        return o;
    }
};

// VEC elements moved by one 16-byte (VEC > 1) or one element-sized (VEC == 1) memory request
template <typename E, int VEC>
struct alignas(sizeof(E) * VEC) Pack {
    E e[VEC];
};

// Per-CTA state in shared memory: the current piece and its scalars. Keeping them out of registers is what lets six
// CTAs (1536 threads) stay resident per SM -- an HBM-bound kernel lives on the number of independent requests in flight.
struct Shared {
    DevPiece p;
    DevScalars s;
};

struct Tile {
    int r0, c0, nr, nc;  // origin and extent of the tile inside S
};

template <typename Ops>
__device__ __forceinline__ typename Ops::E finish(const Shared& sh, typename Ops::E x, const typename Ops::E* d) {
    const unsigned flags = sh.p.flags;
    if (flags & PIECE_CONJ) x = Ops::conj(x);
    if (flags & PIECE_SCALE_ONLY) x = Ops::zero();
    if (!(flags & PIECE_IDENTITY)) x = Ops::axpby(sh.s, x, (flags & PIECE_READ_DST) ? *d : x, flags & PIECE_READ_DST);
    return x;
}

// D = S: requests run along the rows of a column (the contiguous dimension of both sides); a warp request covers
// 32 * VEC * sizeof(E) contiguous bytes (512 B when vectorised).
template <typename Ops, int VEC, int TILE_ROWS, int TC>
__device__ __forceinline__ void copy_tile(const Shared& sh, const Tile& t) {
    using E = typename Ops::E;
    using P = Pack<E, VEC>;
    constexpr int VR = TILE_ROWS / VEC;       // requests per tile column
    constexpr int TOTAL = VR * TC / THREADS;  // requests per thread
    static_assert(TOTAL % CHUNK == 0 || TOTAL < CHUNK, "tile shape");
    constexpr int STEP = TOTAL < CHUNK ? TOTAL : CHUNK;
    const E* __restrict__ src = reinterpret_cast<const E*>(sh.p.src);
    E* __restrict__ dst = reinterpret_cast<E*>(sh.p.dst);
    const long long sld = sh.p.src_ld, dld = sh.p.dst_ld;
    const bool scale_only = sh.p.flags & PIECE_SCALE_ONLY, read_dst = sh.p.flags & PIECE_READ_DST;
#pragma unroll 1
    for (int base = 0; base < TOTAL; base += STEP) {
        P v[STEP];
#pragma unroll
        for (int q = 0; q < STEP; ++q) {
            const int idx = threadIdx.x + (base + q) * THREADS;
            const int r = (idx % VR) * VEC, c = idx / VR;
            if (c < t.nc && r < t.nr && !scale_only) v[q] = *reinterpret_cast<const P*>(src + (t.r0 + r) + (t.c0 + c) * sld);
        }
#pragma unroll
        for (int q = 0; q < STEP; ++q) {
            const int idx = threadIdx.x + (base + q) * THREADS;
            const int r = (idx % VR) * VEC, c = idx / VR;
            if (c < t.nc && r < t.nr) {
                E* d = dst + (t.r0 + r) + (t.c0 + c) * dld;
                P o;
                if (read_dst) o = *reinterpret_cast<const P*>(d);
#pragma unroll
                for (int e = 0; e < VEC; ++e) o.e[e] = finish<Ops>(sh, v[q].e[e], &o.e[e]);
                *reinterpret_cast<P*>(d) = o;
            }
        }
    }
}

// D = S^T without staging through shared memory: every thread owns VEC x VEC micro-tiles, read as VEC requests along S
// rows, transposed in registers and written as VEC requests along D rows. Lanes form an LR x (32/LR) grid of
// micro-tiles, so a warp reads LR * VEC * sizeof(E) contiguous bytes per S column and writes (32/LR) * VEC * sizeof(E)
// contiguous bytes per D column -- whole 32-byte sectors on both sides.
template <typename Ops, int VEC, int TILE_ROWS, int TC>
__device__ __forceinline__ void transpose_tile(const Shared& sh, const Tile& t, int lr_shift) {
    using E = typename Ops::E;
    using P = Pack<E, VEC>;
    constexpr int MR = TILE_ROWS / VEC, MC = TC / VEC;     // micro-tile grid of the CTA tile
    constexpr int BLOCKS_PER_WARP = MR * MC / 32 / WARPS;  // lane blocks (32 micro-tiles each) per warp
    constexpr int PER_ITER = CHUNK / VEC > 0 ? CHUNK / VEC : 1;
    constexpr int ITER = PER_ITER < BLOCKS_PER_WARP ? PER_ITER : BLOCKS_PER_WARP;
    static_assert(BLOCKS_PER_WARP >= 1 && BLOCKS_PER_WARP % ITER == 0, "tile shape");
    const int lc_count = 32 >> lr_shift, BR = MR >> lr_shift;
    const E* __restrict__ src = reinterpret_cast<const E*>(sh.p.src);
    E* __restrict__ dst = reinterpret_cast<E*>(sh.p.dst);
    const long long sld = sh.p.src_ld, dld = sh.p.dst_ld;
    const bool read_dst = sh.p.flags & PIECE_READ_DST;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lr = lane & ((1 << lr_shift) - 1), lc = lane >> lr_shift;
#pragma unroll 1
    for (int base = 0; base < BLOCKS_PER_WARP; base += ITER) {
        P v[ITER][VEC];
#pragma unroll
        for (int q = 0; q < ITER; ++q) {
            const int b = warp + (base + q) * WARPS;
            const int r = (((b % BR) << lr_shift) + lr) * VEC, c = ((b / BR) * lc_count + lc) * VEC;
#pragma unroll
            for (int qc = 0; qc < VEC; ++qc)
                if (r < t.nr && c < t.nc) v[q][qc] = *reinterpret_cast<const P*>(src + (t.r0 + r) + (t.c0 + c + qc) * sld);
        }
#pragma unroll
        for (int q = 0; q < ITER; ++q) {
            const int b = warp + (base + q) * WARPS;
            const int r = (((b % BR) << lr_shift) + lr) * VEC, c = ((b / BR) * lc_count + lc) * VEC;
#pragma unroll
            for (int pr = 0; pr < VEC; ++pr) {
                if (r < t.nr && c < t.nc) {
                    // D(c .. c+VEC-1, r + pr) = S(r + pr, c .. c+VEC-1)
                    E* d = dst + (t.c0 + c) + (t.r0 + r + pr) * dld;
                    P o;
                    if (read_dst) o = *reinterpret_cast<const P*>(d);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) o.e[e] = finish<Ops>(sh, v[q][e].e[pr], &o.e[e]);
                    *reinterpret_cast<P*>(d) = o;
                }
            }
        }
    }
}

// D = S^T staged through a padded shared-memory tile (COSMA_B200_RELAYOUT_SMEM=ON, opt-in until measured): the whole
// CTA reads the tile along S columns and writes it along D columns, element-sized requests, so a warp request covers 32 consecutive
// elements on BOTH sides (512 / 256 / 128 contiguous bytes for 16 / 8 / 4-byte elements) where the register transpose above moves
// 64-128 byte runs. Every thread has 4 / 8 / 8 independent loads in flight (16 / 8 / 4-byte elements) before it stores to shared memory. The row pitch of
// TILE_ROWS + 1 elements makes the transposed reads conflict-free for all three element sizes. Handles ragged edges by itself.
template <typename Ops, int TILE_ROWS, int TC>
__device__ __forceinline__ void transpose_tile_smem(const Shared& sh, const Tile& t, typename Ops::E* tile) {
    using E = typename Ops::E;
    constexpr int PITCH = TILE_ROWS + 1;
    constexpr int PER_THREAD = TILE_ROWS * TC / THREADS;
    static_assert(TILE_ROWS * TC % THREADS == 0 && THREADS % TILE_ROWS == 0 && THREADS % TC == 0, "tile shape");
    const E* __restrict__ src = reinterpret_cast<const E*>(sh.p.src);
    E* __restrict__ dst = reinterpret_cast<E*>(sh.p.dst);
    const long long sld = sh.p.src_ld, dld = sh.p.dst_ld;
    __syncthreads();  // the previous tile has left shared memory
    constexpr int BATCH = PER_THREAD > 8 ? 8 : PER_THREAD;  // loads in flight per thread (more would spill at 40 registers)
#pragma unroll 1
    for (int base = 0; base < PER_THREAD; base += BATCH) {
        E v[BATCH];
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            const int idx = threadIdx.x + (base + q) * THREADS;
            const int r = idx % TILE_ROWS, c = idx / TILE_ROWS;
            if (r < t.nr && c < t.nc) v[q] = src[(t.r0 + r) + (t.c0 + c) * sld];
        }
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            const int idx = threadIdx.x + (base + q) * THREADS;
            const int r = idx % TILE_ROWS, c = idx / TILE_ROWS;
            if (r < t.nr && c < t.nc) tile[c * PITCH + r] = v[q];
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PER_THREAD; ++q) {
        const int idx = threadIdx.x + q * THREADS;
        const int c = idx % TC, r = idx / TC;  // D(c, r) = S(r, c): consecutive lanes walk a D column
        if (r < t.nr && c < t.nc) {
            E* d = dst + (t.c0 + c) + (t.r0 + r) * dld;
            *d = finish<Ops>(sh, tile[c * PITCH + r], d);
        }
    }
}

// Edge tiles whose extent is not a whole number of 16-byte vectors: plain element loop (rare, small).
template <typename Ops, bool TRANSPOSE>
__device__ __forceinline__ void ragged_tile(const Shared& sh, const Tile& t) {
    using E = typename Ops::E;
    const E* __restrict__ src = reinterpret_cast<const E*>(sh.p.src);
    E* __restrict__ dst = reinterpret_cast<E*>(sh.p.dst);
    const long long sld = sh.p.src_ld, dld = sh.p.dst_ld;
    const int n = t.nr * t.nc;
    for (int idx = threadIdx.x; idx < n; idx += THREADS) {
        const int r = idx % t.nr, c = idx / t.nr;
        E x = Ops::zero();
        if (!(sh.p.flags & PIECE_SCALE_ONLY)) x = src[(t.r0 + r) + (t.c0 + c) * sld];
        E* d = TRANSPOSE ? dst + (t.c0 + c) + (t.r0 + r) * dld : dst + (t.r0 + r) + (t.c0 + c) * dld;
        *d = finish<Ops>(sh, x, d);
    }
}

template <typename Ops, bool TRANSPOSE, int VEC, bool SMEM = false>
__global__ void __launch_bounds__(THREADS, 6) relayout_kernel(const DevPiece* __restrict__ pieces, const DevScalars* __restrict__ scalars,
                                                               int n_pieces, long long total_tiles, long long tiles_per_cta, int lr_shift) {
    using E = typename Ops::E;
    constexpr int TILE_ROWS = relayout_tile_rows(sizeof(E), TRANSPOSE), TC = relayout_tile_cols(sizeof(E), TRANSPOSE);
    __shared__ Shared sh;
    __shared__ alignas(16) unsigned char tile_mem[SMEM ? (TILE_ROWS + 1) * TC * sizeof(E) : 16];
    long long t = static_cast<long long>(blockIdx.x) * tiles_per_cta;
    const long long t_end = min(total_tiles, t + tiles_per_cta);
    if (t >= t_end) return;

    // the piece containing tile t: last piece with tile_begin <= t (every thread searches: uniform, L2-resident)
    int lo = 0, hi = n_pieces - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (pieces[mid].tile_begin <= t) lo = mid; else hi = mid - 1;
    }
    int pi = lo - 1;
    long long p_begin = 0, p_end = t;  // forces a descriptor load on the first iteration

    for (; t < t_end; ++t) {
        if (t >= p_end) {  // uniform across the CTA
            do {
                ++pi;
                p_begin = pieces[pi].tile_begin;
                p_end = pi + 1 < n_pieces ? pieces[pi + 1].tile_begin : total_tiles;
            } while (t >= p_end);
            __syncthreads();  // everyone is done with the previous piece
            if (threadIdx.x < sizeof(DevPiece) / 4)
                reinterpret_cast<unsigned*>(&sh.p)[threadIdx.x] = reinterpret_cast<const unsigned*>(pieces + pi)[threadIdx.x];
            const int param = pieces[pi].param;
            if (threadIdx.x >= 32 && threadIdx.x < 32 + sizeof(DevScalars) / 4) {
                const unsigned w = threadIdx.x - 32;
                unsigned val = 0;
                if (param >= 0) val = reinterpret_cast<const unsigned*>(scalars + param)[w];
                reinterpret_cast<unsigned*>(&sh.s)[w] = val;
            }
            __syncthreads();
        }
        const int lt = static_cast<int>(t - p_begin);
        const int tiles_r = (sh.p.rows + TILE_ROWS - 1) / TILE_ROWS;
        Tile tc;
        tc.r0 = (lt % tiles_r) * TILE_ROWS;
        tc.c0 = (lt / tiles_r) * TC;
        tc.nr = min(TILE_ROWS, sh.p.rows - tc.r0);
        tc.nc = min(TC, sh.p.cols - tc.c0);
        if (SMEM) transpose_tile_smem<Ops, TILE_ROWS, TC>(sh, tc, reinterpret_cast<E*>(tile_mem));
        else if (VEC > 1 && (tc.nr % VEC != 0 || tc.nc % VEC != 0)) ragged_tile<Ops, TRANSPOSE>(sh, tc);
        else if (TRANSPOSE) transpose_tile<Ops, VEC, TILE_ROWS, TC>(sh, tc, lr_shift);
        else copy_tile<Ops, VEC, TILE_ROWS, TC>(sh, tc);
    }
}

// Tunables (defaults chosen from measurements on B200, profiles/; overridable for experiments):
//   COSMA_B200_RELAYOUT_LR = 4 | 8 | 16   rows of the lane grid used by transposes
struct Tuning {
    // transposes through a padded shared-memory tile (full-line requests on both global sides) or through registers (lane-grid
    // shuffle). Measured on B200 (profiles/r2_relayout_sweep*.txt, 16384^2 in 256^2 pieces, fraction of the 6543 GB/s copy peak):
    //   shared:   z 0.85-0.86   d 0.82   c 0.81   s 0.57          registers:   z 0.67   d 0.73   c 0.76   s 0.73
    // -> default (-1): shared memory for elements of 8 bytes and more, registers for 4-byte ones. COSMA_B200_RELAYOUT_SMEM=ON|OFF forces one.
    int smem_transpose = -1;
    int lr_shift = 2;  // 4 x 8 lane grid: 64-byte read runs, 128-byte write runs (best of 4|8|16 on B200, profiles/r1_relayout_sweep.txt)
};
const Tuning& tuning() {
    static Tuning t = [] {
        Tuning v;
        if (const char* e = std::getenv("COSMA_B200_RELAYOUT_SMEM"))
            if (e[0] == 'O') v.smem_transpose = e[1] == 'N' ? 1 : 0;
        if (const char* e = std::getenv("COSMA_B200_RELAYOUT_LR")) {
            const int lr = std::atoi(e);
            v.lr_shift = lr == 8 ? 3 : lr == 16 ? 4 : lr == 2 ? 1 : 2;
        }
        return v;
    }();
    return t;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

}  // namespace

void relayout_normalise(const std::vector<costa::piece>& pieces, const char* src_base, char* dst_base, int elem_bytes,
                        const std::vector<costa::transform_spec>& specs, RelayoutHostList& out) {
    out.scalars.clear();
    for (const auto& s : specs) out.scalars.push_back(DevScalars{{s.alpha[0], s.alpha[1]}, {s.beta[0], s.beta[1]}});
    for (const auto& p : pieces) {
        if (p.n_rows <= 0 || p.n_cols <= 0) continue;
        DevPiece d{};
        d.src = (src_base ? src_base : static_cast<const char*>(nullptr)) + reinterpret_cast<std::intptr_t>(p.src);
        d.dst = (dst_base ? dst_base : static_cast<char*>(nullptr)) + reinterpret_cast<std::intptr_t>(p.dst);
        const bool s_row = p.src_ordering == 'R', d_row = p.dst_ordering == 'R';
        d.rows = s_row ? p.n_cols : p.n_rows;
        d.cols = s_row ? p.n_rows : p.n_cols;
        d.src_ld = p.src_ld;
        d.dst_ld = p.dst_ld;
        const bool will_transpose = !p.scale_only && ((p.transpose ? 1 : 0) ^ (s_row ? 1 : 0) ^ (d_row ? 1 : 0));
        d.flags = (will_transpose ? PIECE_TRANSPOSE : 0u) | (p.conjugate ? PIECE_CONJ : 0u) | (p.scale_only ? PIECE_SCALE_ONLY : 0u);
        d.param = p.transform;
        bool identity = true, rd = false;
        if (p.transform >= 0) {
            const auto& s = specs[p.transform];
            identity = s.alpha[0] == 1.0 && s.alpha[1] == 0.0 && s.beta[0] == 0.0 && s.beta[1] == 0.0;
            rd = s.beta[0] != 0.0 || s.beta[1] != 0.0;
        }
        if (p.scale_only) identity = false;
        if (identity) d.flags |= PIECE_IDENTITY;
        if (rd) { d.flags |= PIECE_READ_DST; out.reads_dst = true; }
        // 16-byte requests are legal when both sides keep 16-byte alignment in every column
        const auto aligned16 = [&](const void* ptr, std::int64_t ld) {
            return (reinterpret_cast<std::uintptr_t>(ptr) & 15) == 0 && (ld * elem_bytes) % 16 == 0;
        };
        const bool vec = elem_bytes < 16 ? (aligned16(d.src, d.src_ld) && aligned16(d.dst, d.dst_ld)) : true;
        if (vec) d.flags |= PIECE_VEC16;
        const int c = (will_transpose ? 2 : 0) | (vec ? 1 : 0);
        d.tile_begin = out.tiles[c];
        const int tile_rows = relayout_tile_rows(elem_bytes, will_transpose), tile_cols = relayout_tile_cols(elem_bytes, will_transpose);
        out.tiles[c] += static_cast<std::int64_t>((d.rows + tile_rows - 1) / tile_rows) * ((d.cols + tile_cols - 1) / tile_cols);
        out.elements += static_cast<std::int64_t>(d.rows) * d.cols;
        out.cls[c].push_back(d);
    }
}

int relayout_upload(const RelayoutHostList& list, RelayoutBatch& out) {
    out.elements = list.elements;
    out.reads_dst = list.reads_dst;
    bool any = false;
    for (int c = 0; c < RELAYOUT_CLASSES; ++c) {
        out.n_pieces[c] = static_cast<int>(list.cls[c].size());
        out.tiles[c] = list.tiles[c];
        if (list.cls[c].empty()) continue;
        any = true;
        const size_t bytes = list.cls[c].size() * sizeof(DevPiece);
        if (cudaMalloc(reinterpret_cast<void**>(&out.d_pieces[c]), bytes) != cudaSuccess) {
            set_last_error("relayout: cudaMalloc of the piece list failed");
            return COSMA_B200_OUT_OF_MEMORY;
        }
        if (cudaMemcpy(out.d_pieces[c], list.cls[c].data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_last_error("relayout: upload of the piece list failed");
            return COSMA_B200_CUDA_ERROR;
        }
    }
    if (!any) return COSMA_B200_OK;
    const size_t sbytes = std::max<size_t>(list.scalars.size(), 1) * sizeof(DevScalars);
    if (cudaMalloc(reinterpret_cast<void**>(&out.d_scalars), sbytes) != cudaSuccess) {
        set_last_error("relayout: cudaMalloc of the scalar table failed");
        return COSMA_B200_OUT_OF_MEMORY;
    }
    if (!list.scalars.empty() &&
        cudaMemcpy(out.d_scalars, list.scalars.data(), list.scalars.size() * sizeof(DevScalars), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_last_error("relayout: upload of the scalar table failed");
        return COSMA_B200_CUDA_ERROR;
    }
    return COSMA_B200_OK;
}

void relayout_free(RelayoutBatch& b) {
    for (auto& p : b.d_pieces)
        if (p) cudaFree(p);
    if (b.d_scalars) cudaFree(b.d_scalars);
    b = RelayoutBatch{};
}

namespace {
template <typename Ops, bool TRANSPOSE, int VEC, bool SMEM = false>
int launch_class(const DevPiece* pieces, int n_pieces, std::int64_t tiles, const DevScalars* scalars, cudaStream_t stream) {
    // grid = SMs x resident CTAs (no partial wave); every CTA walks a contiguous range of tiles
    static int resident = 0;
    if (resident == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, relayout_kernel<Ops, TRANSPOSE, VEC, SMEM>, THREADS, 0) != cudaSuccess || resident <= 0) {
            cudaGetLastError();
            resident = 6;
        }
    }
    const long long max_ctas = static_cast<long long>(sm_count()) * resident;
    long long ctas = std::min<long long>(tiles, max_ctas);
    const long long per = (tiles + ctas - 1) / ctas;
    ctas = (tiles + per - 1) / per;
    relayout_kernel<Ops, TRANSPOSE, VEC, SMEM><<<dim3(static_cast<unsigned>(ctas)), dim3(THREADS), 0, stream>>>(pieces, scalars, n_pieces, tiles, per, tuning().lr_shift);
    if (cudaGetLastError() != cudaSuccess) {
        set_last_error("relayout: kernel launch failed");
        return COSMA_B200_CUDA_ERROR;
    }
    return COSMA_B200_OK;
}

template <typename Ops>
int launch_typed(const RelayoutBatch& b, cudaStream_t stream, int* launches) {
    constexpr int VEC = 16 / static_cast<int>(sizeof(typename Ops::E));
    const bool smem = tuning().smem_transpose < 0 ? sizeof(typename Ops::E) >= 8 : tuning().smem_transpose == 1;
    for (int c = 0; c < RELAYOUT_CLASSES; ++c) {
        if (b.tiles[c] == 0) continue;
        int st;
        switch (c) {
            case 0: st = launch_class<Ops, false, 1>(b.d_pieces[c], b.n_pieces[c], b.tiles[c], b.d_scalars, stream); break;
            case 1: st = launch_class<Ops, false, VEC>(b.d_pieces[c], b.n_pieces[c], b.tiles[c], b.d_scalars, stream); break;
            case 2:
                st = smem ? launch_class<Ops, true, 1, true>(b.d_pieces[c], b.n_pieces[c], b.tiles[c], b.d_scalars, stream)
                                             : launch_class<Ops, true, 1>(b.d_pieces[c], b.n_pieces[c], b.tiles[c], b.d_scalars, stream);
                break;
            default:
                st = smem ? launch_class<Ops, true, VEC, true>(b.d_pieces[c], b.n_pieces[c], b.tiles[c], b.d_scalars, stream)
                                             : launch_class<Ops, true, VEC>(b.d_pieces[c], b.n_pieces[c], b.tiles[c], b.d_scalars, stream);
                break;
        }
        if (st != COSMA_B200_OK) return st;
        if (launches) ++*launches;
    }
    return COSMA_B200_OK;
}
}  // namespace

int relayout_launch(const RelayoutBatch& b, char dtype, cudaStream_t stream, int* launches) {
    if (b.empty()) return COSMA_B200_OK;
    switch (dtype) {
        case 's': return launch_typed<RealOps<float>>(b, stream, launches);
        case 'd': return launch_typed<RealOps<double>>(b, stream, launches);
        case 'c': return launch_typed<CplxOps<float, float2>>(b, stream, launches);
        case 'z': return launch_typed<CplxOps<double, double2>>(b, stream, launches);
        default:
            set_last_error("relayout: dtype must be one of s, d, c, z");
            return COSMA_B200_INVALID_ARG;
    }
}

}  // namespace cosma_b200
