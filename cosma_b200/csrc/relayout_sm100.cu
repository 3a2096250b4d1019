// R3/R4: batched block copy / transpose (+conjugate, +alpha/beta) for sm_100a -- the device form of COSTA's
// copy_and_transform (reference libs/COSTA/src/costa/grid2grid/memory_utils.hpp:287-346: copy2D :47-85,
// transpose_col_major :88-166, transpose_row_major :169-250), which the reference runs on the host with OpenMP for
// every message of a layout transformation (communication_data.cpp:166-302).
//
// One launch handles a whole list of pieces. HBM-bound: 2 * elements * sizeof(T) algorithmic bytes per launch
// (3 * when beta != 0). Design:
//  * every piece is normalised on the host to a column-major view (row-major storage = the transposed column-major
//    matrix), so the four ordering combinations of the reference collapse into "copy" or "transpose";
//  * pieces are cut into 32 x 32-element tiles, numbered globally through a prefix sum kept in the descriptors;
//    each CTA of a grid sized in multiples of the SM count takes a CONTIGUOUS range of tiles (one binary search, then
//    a linear walk), column-major inside a piece so consecutive tiles touch neighbouring DRAM pages;
//  * copies read and write along the contiguous dimension (32 lanes x sizeof(T) per request, four requests in flight
//    per warp before the first store); transposes stage the tile through padded shared memory so both the global read
//    and the global write stay coalesced;
//  * alpha = 1, beta = 0 moves bits only (conjugation flips one sign bit) -> bit-exact; otherwise
//    dst = beta*dst + alpha*op(src) is evaluated with unfused multiplies and adds in the reference's order, and
//    beta == 0 never reads dst.
#include "relayout_sm100.h"
#include "../../include/cosma_b200.h"

#include <algorithm>
#include <string>

namespace cosma_b200 {
void set_last_error(const std::string& msg);

namespace {

constexpr int T = RELAYOUT_TILE;
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int PER_WARP = T / WARPS;  // tile columns handled by one warp

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

template <typename R>
struct Scal {
    R ar, ai, br, bi;
};

// real element
template <typename R>
struct RealOps {
    using E = R;
    static __device__ __forceinline__ E conj(E v) { return v; }
    static __device__ __forceinline__ E axpby(const Scal<R>& s, E x, E d, bool read_dst) {
        const R ax = mul_rn(s.ar, x);
        return read_dst ? add_rn(mul_rn(s.br, d), ax) : ax;
    }
};
template <typename R, typename R2>
struct CplxOps {
    using E = R2;
    static __device__ __forceinline__ E conj(E v) { v.y = -v.y; return v; }
    static __device__ __forceinline__ E mul(R ar, R ai, E x) {
        E o;
        o.x = sub_rn(mul_rn(ar, x.x), mul_rn(ai, x.y));
        o.y = add_rn(mul_rn(ar, x.y), mul_rn(ai, x.x));
        return o;
    }
    static __device__ __forceinline__ E axpby(const Scal<R>& s, E x, E d, bool read_dst) {
        const E ax = mul(s.ar, s.ai, x);
        if (!read_dst) return ax;
        const E bd = mul(s.br, s.bi, d);
        E o;
        o.x = add_rn(bd.x, ax.x);
        o.y = add_rn(bd.y, ax.y);
        return o;
    }
};

template <typename R, typename Ops>
__global__ void __launch_bounds__(THREADS) relayout_kernel(const DevPiece* __restrict__ pieces, const DevScalars* __restrict__ scalars,
                                                            int n_pieces, long long total_tiles, long long tiles_per_cta) {
    using E = typename Ops::E;
    __shared__ E tile[T][T + 1];
    long long t = static_cast<long long>(blockIdx.x) * tiles_per_cta;
    const long long t_end = min(total_tiles, t + tiles_per_cta);
    if (t >= t_end) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

    // the piece containing tile t: last piece with tile_begin <= t
    int lo = 0, hi = n_pieces - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (pieces[mid].tile_begin <= t) lo = mid; else hi = mid - 1;
    }
    int pi = lo;
    DevPiece p = pieces[pi];
    long long p_end = pi + 1 < n_pieces ? pieces[pi + 1].tile_begin : total_tiles;
    Scal<R> sc{R(1), R(0), R(0), R(0)};
    auto load_scalars = [&]() {
        if (p.param >= 0) {
            const DevScalars s = scalars[p.param];
            sc = Scal<R>{static_cast<R>(s.alpha[0]), static_cast<R>(s.alpha[1]), static_cast<R>(s.beta[0]), static_cast<R>(s.beta[1])};
        }
    };
    load_scalars();

    for (; t < t_end; ++t) {
        while (t >= p_end) {
            ++pi;
            p = pieces[pi];
            p_end = pi + 1 < n_pieces ? pieces[pi + 1].tile_begin : total_tiles;
            load_scalars();
        }
        const long long lt = t - p.tile_begin;
        const int tiles_r = (p.rows + T - 1) / T;
        const int ti = static_cast<int>(lt % tiles_r), tj = static_cast<int>(lt / tiles_r);
        const int r0 = ti * T, c0 = tj * T;
        const int nr = min(T, p.rows - r0), nc = min(T, p.cols - c0);
        const E* __restrict__ src = reinterpret_cast<const E*>(p.src);
        E* __restrict__ dst = reinterpret_cast<E*>(p.dst);
        const bool conj = p.flags & PIECE_CONJ, identity = p.flags & PIECE_IDENTITY, read_dst = p.flags & PIECE_READ_DST;
        const bool scale_only = p.flags & PIECE_SCALE_ONLY;  // dst = beta*dst (copy path only; alpha is forced to 0 on the host)

        if (!(p.flags & PIECE_TRANSPOSE)) {
            E v[PER_WARP];
#pragma unroll
            for (int q = 0; q < PER_WARP; ++q) {
                const int c = w + q * WARPS;
                if (lane < nr && c < nc && !scale_only) v[q] = src[(r0 + lane) + static_cast<long long>(c0 + c) * p.src_ld];
            }
#pragma unroll
            for (int q = 0; q < PER_WARP; ++q) {
                const int c = w + q * WARPS;
                if (lane < nr && c < nc) {
                    E* d = dst + (r0 + lane) + static_cast<long long>(c0 + c) * p.dst_ld;
                    E x = conj ? Ops::conj(v[q]) : v[q];
                    if (scale_only) x = E{};
                    if (!identity) x = Ops::axpby(sc, x, read_dst ? *d : x, read_dst);
                    *d = x;
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < PER_WARP; ++q) {
                const int c = w + q * WARPS;
                if (lane < nr && c < nc) {
                    const E x = src[(r0 + lane) + static_cast<long long>(c0 + c) * p.src_ld];
                    tile[c][lane] = conj ? Ops::conj(x) : x;
                }
            }
            __syncthreads();
            // D(c0 + j, r0 + i) = S(r0 + i, c0 + j): lanes run along j, the contiguous dimension of D
#pragma unroll
            for (int q = 0; q < PER_WARP; ++q) {
                const int i = w + q * WARPS;
                if (lane < nc && i < nr) {
                    E* d = dst + (c0 + lane) + static_cast<long long>(r0 + i) * p.dst_ld;
                    E x = tile[lane][i];
                    if (!identity) x = Ops::axpby(sc, x, read_dst ? *d : x, read_dst);
                    *d = x;
                }
            }
            __syncthreads();
        }
    }
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
                        const std::vector<costa::transform_spec>& specs, std::vector<DevPiece>& out, std::vector<DevScalars>& scalars,
                        std::int64_t* total_tiles, std::int64_t* elements, bool* reads_dst) {
    (void)elem_bytes;
    scalars.clear();
    for (const auto& s : specs) scalars.push_back(DevScalars{{s.alpha[0], s.alpha[1]}, {s.beta[0], s.beta[1]}});
    std::int64_t tiles = *total_tiles, elems = *elements;
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
        if (rd) { d.flags |= PIECE_READ_DST; *reads_dst = true; }
        d.tile_begin = tiles;
        tiles += static_cast<std::int64_t>((d.rows + T - 1) / T) * ((d.cols + T - 1) / T);
        elems += static_cast<std::int64_t>(d.rows) * d.cols;
        out.push_back(d);
    }
    *total_tiles = tiles;
    *elements = elems;
}

int relayout_upload(const std::vector<DevPiece>& pieces, const std::vector<DevScalars>& scalars, RelayoutBatch& out) {
    out.n_pieces = static_cast<int>(pieces.size());
    if (pieces.empty()) return COSMA_B200_OK;
    if (cudaMalloc(reinterpret_cast<void**>(&out.d_pieces), pieces.size() * sizeof(DevPiece)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&out.d_scalars), std::max<size_t>(scalars.size(), 1) * sizeof(DevScalars)) != cudaSuccess) {
        set_last_error("relayout: cudaMalloc of the piece list failed");
        return COSMA_B200_OUT_OF_MEMORY;
    }
    if (cudaMemcpy(out.d_pieces, pieces.data(), pieces.size() * sizeof(DevPiece), cudaMemcpyHostToDevice) != cudaSuccess ||
        (!scalars.empty() &&
         cudaMemcpy(out.d_scalars, scalars.data(), scalars.size() * sizeof(DevScalars), cudaMemcpyHostToDevice) != cudaSuccess)) {
        set_last_error("relayout: upload of the piece list failed");
        return COSMA_B200_CUDA_ERROR;
    }
    return COSMA_B200_OK;
}

void relayout_free(RelayoutBatch& b) {
    if (b.d_pieces) cudaFree(b.d_pieces);
    if (b.d_scalars) cudaFree(b.d_scalars);
    b = RelayoutBatch{};
}

int relayout_launch(const RelayoutBatch& b, char dtype, cudaStream_t stream) {
    if (b.n_pieces == 0 || b.total_tiles == 0) return COSMA_B200_OK;
    // grid: a multiple of the SM count, at most 8 resident CTAs per SM; every CTA walks a contiguous tile range
    const long long max_ctas = static_cast<long long>(sm_count()) * 8;
    long long ctas = std::min<long long>(b.total_tiles, max_ctas);
    const long long per = (b.total_tiles + ctas - 1) / ctas;
    ctas = (b.total_tiles + per - 1) / per;
    const dim3 grid(static_cast<unsigned>(ctas)), block(THREADS);
    switch (dtype) {
        case 's': relayout_kernel<float, RealOps<float>><<<grid, block, 0, stream>>>(b.d_pieces, b.d_scalars, b.n_pieces, b.total_tiles, per); break;
        case 'd': relayout_kernel<double, RealOps<double>><<<grid, block, 0, stream>>>(b.d_pieces, b.d_scalars, b.n_pieces, b.total_tiles, per); break;
        case 'c': relayout_kernel<float, CplxOps<float, float2>><<<grid, block, 0, stream>>>(b.d_pieces, b.d_scalars, b.n_pieces, b.total_tiles, per); break;
        case 'z': relayout_kernel<double, CplxOps<double, double2>><<<grid, block, 0, stream>>>(b.d_pieces, b.d_scalars, b.n_pieces, b.total_tiles, per); break;
        default:
            set_last_error("relayout: dtype must be one of s, d, c, z");
            return COSMA_B200_INVALID_ARG;
    }
    if (cudaGetLastError() != cudaSuccess) {
        set_last_error("relayout: kernel launch failed");
        return COSMA_B200_CUDA_ERROR;
    }
    return COSMA_B200_OK;
}

}  // namespace cosma_b200
