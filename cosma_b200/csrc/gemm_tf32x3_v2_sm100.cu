// K3/K4 second generation: FP32 SGEMM / CGEMM on tcgen05 with the 3xTF32 split, the A operand fed through TENSOR MEMORY.
//
// The first generation (gemm_tf32x3_sm100.cu) is bound by shared-memory bandwidth, not by the tensor pipe (ncu, profiles/
// r2_ncu_sgemm8192_full.json: tensor pipe 44.5 % active; per 128 x 128 x 32 block the TMA writes 32 KB, the split stage reads 32 KB and
// writes 64 KB, the MMAs read 80 KB = 208 KB at 128 B/clk = 1660 cycles against 768 cycles of tensor work). Here
//   * the split A tiles (A_hi, A_lo) never touch shared memory: the A-split warps hold one row per thread in registers and write
//     hi / lo with tcgen05.st into a two-stage ring in TMEM; the MMAs take A from TMEM (the "TS" form of tcgen05.mma);
//   * the small cross terms (hi*lo + lo*hi) accumulate in ONE long-lived TMEM accumulator per tile -- their round-toward-zero drift
//     is 2^-11 of that of the large terms -- so only the large hi*hi accumulator is chunked and drained every CHUNK_KB k blocks
//     (round-to-nearest adds in registers, see "Numerics" in gemm_tf32x3_sm100.cu), which halves the drain traffic;
//   * shared memory then carries per block: TMA 32 KB + split reads 32 KB + B_hi/B_lo writes 32 KB + MMA reads of B 48 KB = 144 KB
//     = 1125 cycles: the tensor pipe can be ~68 % busy instead of ~46 %.
// TMEM (512 columns): [0,128) big accumulator 0 | [128,256) big accumulator 1 | [256,384) small accumulator | [384,512) A ring
// (2 stages x {A_hi 32 columns, A_lo 32 columns}; lane = row of the tile, column = k).
// 16 warps in four warpgroups with re-balanced registers (setmaxnreg): drain + epilogue (208) | A split -> TMEM (128) |
// B split -> shared memory (80) | TMA producer + MMA issuer (40).
// CGEMM: all nine op(A), op(B) combinations run here -- the split stages build the real embedding [[Ar,-Ai],[Ai,Ar]] of op(A) and the
// real view of op(B) from either storage order, conjugation is a sign flip on the way.
#include "sm100_common.cuh"
#include "gemm_tf32x3_sm100.h"
#include "cta_budget.h"

#include <mutex>
#include <string>

namespace cosma_b200 {
void set_last_error(const std::string& msg);

namespace {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 32;  // floats: one 128-byte swizzle row
constexpr int RAW_STAGES = 4;
constexpr int OP_STAGES = 2;
constexpr int TILE_BYTES = 128 * BK * 4;          // 16 KB: a 128 x 32 FP32 operand tile
constexpr int RAW_STAGE_BYTES = 2 * TILE_BYTES;   // A raw, B raw
constexpr int BOP_STAGE_BYTES = 2 * TILE_BYTES;   // B_hi, B_lo
constexpr int SMEM_BYTES = RAW_STAGES * RAW_STAGE_BYTES + OP_STAGES * BOP_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int THREADS = 512;                      // 4 warpgroups
constexpr int SPLIT_THREADS = 128;                // per operand
constexpr int CHUNK_KB = 2;                       // k blocks of hi*hi chained inside the tensor core before the drain
constexpr int TMEM_COLS = 512;
constexpr uint32_t COL_BIG = 0, COL_SMALL = 256, COL_A = 384;

enum AMode : int { A_KMAJOR = 0, A_ROWMAJOR = 1, A_CPLX_ROW = 2, A_CPLX_K = 3 };
// A_KMAJOR  : k contiguous (op(A) = A^T, real): box {32 k, 128 rows}, 128B swizzle
// A_ROWMAJOR: rows contiguous (op(A) = A, real): box {128 rows, 32 k}
// A_CPLX_ROW: complex, op(A) = A: real view rows contiguous, box {128 real rows, 16 complex k}
// A_CPLX_K  : complex, op(A) = A^T | A^H: real view k contiguous, box {32 real k, 64 complex rows}, 128B swizzle
enum BMode : int { B_KMAJOR = 0, B_ROWMAJOR = 1, B_CPLX_ROW = 2 };
// B_KMAJOR  : k contiguous (op(B) = B; complex: the real view of B IS the operand): box {32 k, 128 columns}, 128B swizzle
// B_ROWMAJOR: columns contiguous (op(B) = B^T, real): box {128 columns, 32 k}
// B_CPLX_ROW: complex, op(B) = B^T | B^H: real view columns contiguous, box {256 reals = 128 complex columns, 16 complex k}

struct Params {
    int64_t m, n, k;  // extents of the REAL problem the tensor cores see (CGEMM: 2m, n, 2k)
    float alpha[2], beta[2];
    float* C;
    int64_t ldc;      // in floats of the real view
    int tiles_m, tiles_n, num_kb;
    int cplx;
    float conj_a, conj_b;  // -1 where the operand is conjugated, else +1
};

// ---- tcgen05 / TMEM wrappers (raw PTX) ------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]^T, TF32 inputs, FP32 accumulate ("TS" form: A from tensor memory)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// thread i of the warp writes 32 consecutive 32-bit columns of TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled shared-memory operand descriptor (see gemm_tf32x3_sm100.cu)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    return static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// D FP32 (1 << 4), A and B TF32 (2 << 7, 2 << 10), K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);

__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// One row (32 k values) of the A tile for thread = row r, from the raw tile in shared memory.
template <int MODE>
__device__ __forceinline__ void load_a_row(uint32_t raw, int r, float conj, float (&x)[32]) {
    if (MODE == A_ROWMAJOR) {
        // raw[k][r], r contiguous: a warp reads 32 consecutive words per k
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = lds32(raw + static_cast<uint32_t>(j * 128 + r) * 4);
    } else if (MODE == A_KMAJOR) {
        // TMA wrote row r at r * 128 bytes with the 128B swizzle: 16-byte chunk c sits at c ^ (r & 7)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 v = lds128(raw + static_cast<uint32_t>(r) * 128 + static_cast<uint32_t>((c ^ (r & 7)) << 4));
            x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
        }
    } else if (MODE == A_CPLX_ROW) {
        // raw[l][x'] = component (x' & 1) of complex row x' >> 1 at complex k index l (16 per block). Row r = 2i + d of the embedding
        // holds at real k index 2l + c:  d == c ? Ar : (d == 0 ? -Ai : +Ai)
        const float sgn = (r & 1) ? 1.0f : -1.0f;
#pragma unroll
        for (int l = 0; l < 16; ++l) {
            x[2 * l] = lds32(raw + static_cast<uint32_t>(l * 128 + r) * 4);
            x[2 * l + 1] = sgn * lds32(raw + static_cast<uint32_t>(l * 128 + (r ^ 1)) * 4);
        }
    } else {
        // A_CPLX_K: raw row i = r >> 1 holds (Ar_0, Ai_0, Ar_1, Ai_1, ...) of complex row i of op(A) before conjugation, swizzled as above
        const int i = r >> 1;
        const bool d = r & 1;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 v = lds128(raw + static_cast<uint32_t>(i) * 128 + static_cast<uint32_t>((c ^ (i & 7)) << 4));
            const float i0 = conj * v.y, i1 = conj * v.w;
            x[4 * c] = d ? i0 : v.x;
            x[4 * c + 1] = d ? v.x : -i0;
            x[4 * c + 2] = d ? i1 : v.z;
            x[4 * c + 3] = d ? v.z : -i1;
        }
    }
}

// B split: raw tile -> B_hi, B_lo in the canonical K-major 128B-swizzled layout; t = 0..127. Items are (row, 16-byte chunk) pairs,
// 1024 per tile; a warp covers 32 consecutive rows of one chunk (or 32 consecutive vectors), all accesses conflict-free.
template <int MODE>
__device__ __forceinline__ void split_b(uint32_t raw, uint32_t hi, uint32_t lo, int t, float conj) {
#pragma unroll 2
    for (int item = t; item < 1024; item += SPLIT_THREADS) {
        uint32_t h[4], l[4], off;
        if (MODE == B_KMAJOR) {
            off = static_cast<uint32_t>(item) * 16;
            const float4 v = lds128(raw + off);
            split(v.x, h[0], l[0]); split(v.y, h[1], l[1]); split(v.z, h[2], l[2]); split(v.w, h[3], l[3]);
        } else {
            const int r = item & 127, c = item >> 7;
            off = static_cast<uint32_t>(r) * 128 + static_cast<uint32_t>((c ^ (r & 7)) << 4);
            float x[4];
            if (MODE == B_ROWMAJOR) {
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = lds32(raw + static_cast<uint32_t>((4 * c + j) * 128 + r) * 4);
            } else {
                // B_CPLX_ROW: raw[l][2j + comp], 256 floats per l; row j of the operand holds at real k index 2l + comp: Br | conj * Bi
                const float2 v0 = lds64(raw + static_cast<uint32_t>((2 * c) * 256 + 2 * r) * 4);
                const float2 v1 = lds64(raw + static_cast<uint32_t>((2 * c + 1) * 256 + 2 * r) * 4);
                x[0] = v0.x; x[1] = conj * v0.y; x[2] = v1.x; x[3] = conj * v1.y;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) split(x[j], h[j], l[j]);
        }
        sts128(hi + off, h[0], h[1], h[2], h[3]);
        sts128(lo + off, l[0], l[1], l[2], l[3]);
    }
}

__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int& tm, int& tn) {
    constexpr int G = 16;
    const int band_tiles = G * tiles_n;
    const int band = tile / band_tiles;
    const int rem = tile - band * band_tiles;
    const int rows_in_band = min(G, tiles_m - band * G);
    tn = rem / rows_in_band;
    tm = band * G + rem % rows_in_band;
}

template <int MODE_A, int MODE_B>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32x3_v2_sm100_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* aligned = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t raw_ring = base;
    const uint32_t bop_ring = base + RAW_STAGES * RAW_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(aligned + RAW_STAGES * RAW_STAGE_BYTES + OP_STAGES * BOP_STAGE_BYTES);
    uint64_t* raw_full = bars;                      // [RAW_STAGES] TMA -> both split warpgroups
    uint64_t* raw_empty = raw_full + RAW_STAGES;    // [RAW_STAGES] both split warpgroups -> TMA
    uint64_t* a_full = raw_empty + RAW_STAGES;      // [OP_STAGES]  A split (TMEM) -> MMA
    uint64_t* b_full = a_full + OP_STAGES;          // [OP_STAGES]  B split (shared memory) -> MMA
    uint64_t* op_empty = b_full + OP_STAGES;        // [OP_STAGES]  MMA (commit) -> both split warpgroups
    uint64_t* acc_full = op_empty + OP_STAGES;      // [2]          MMA (commit) -> drain: a chunk of hi*hi
    uint64_t* acc_empty = acc_full + 2;             // [2]          drain -> MMA
    uint64_t* small_full = acc_empty + 2;           // [1]          MMA (commit) -> drain: the tile's cross terms
    uint64_t* small_empty = small_full + 1;         // [1]          drain -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(small_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wg = warp >> 2;
    if (threadIdx.x == 0) {
        for (int s = 0; s < RAW_STAGES; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 2 * SPLIT_THREADS); }
        for (int s = 0; s < OP_STAGES; ++s) { mbar_init(&a_full[s], SPLIT_THREADS); mbar_init(&b_full[s], SPLIT_THREADS); mbar_init(&op_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128); }
        mbar_init(small_full, 1);
        mbar_init(small_empty, 128);
        fence_barrier_init();
    }
    if (warp == 13) tmem_alloc(tmem_slot, TMEM_COLS);
    if (warp == 12 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_tiles = p.tiles_m * p.tiles_n;
    constexpr uint32_t A_RAW_BYTES = (MODE_A == A_CPLX_ROW || MODE_A == A_CPLX_K) ? TILE_BYTES / 2 : TILE_BYTES;

    if (wg == 3) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 12 && lane == 0) {
            // ===== TMA producer =====
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int tm, tn;
                tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
                for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                    const uint32_t s = it % RAW_STAGES, ph = (it / RAW_STAGES) & 1;
                    mbar_wait(&raw_empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&raw_full[s], A_RAW_BYTES + TILE_BYTES);
                    uint8_t* a_dst = aligned + s * RAW_STAGE_BYTES;
                    uint8_t* b_dst = a_dst + TILE_BYTES;
                    if (MODE_A == A_KMAJOR) tma_load_2d(a_dst, &map_a, &raw_full[s], kb * BK, tm * BM);
                    else if (MODE_A == A_ROWMAJOR) tma_load_2d(a_dst, &map_a, &raw_full[s], tm * BM, kb * BK);
                    else if (MODE_A == A_CPLX_ROW) tma_load_2d(a_dst, &map_a, &raw_full[s], tm * BM, kb * (BK / 2));
                    else tma_load_2d(a_dst, &map_a, &raw_full[s], kb * BK, tm * (BM / 2));
                    if (MODE_B == B_KMAJOR) tma_load_2d(b_dst, &map_b, &raw_full[s], kb * BK, tn * BN);
                    else if (MODE_B == B_ROWMAJOR) tma_load_2d(b_dst, &map_b, &raw_full[s], tn * BN, kb * BK);
                    else tma_load_2d(b_dst, &map_b, &raw_full[s], tn * (2 * BN), kb * (BK / 2));
                }
            }
        } else if (warp == 13 && lane == 0) {
            // ===== MMA issuer =====
            uint32_t it = 0, chunk = 0, tile_it = 0;
            const uint32_t d_small = tmem_base + COL_SMALL;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
                mbar_wait(small_empty, (tile_it & 1) ^ 1);  // the drain warps have read the previous tile's cross terms
                for (int kb0 = 0; kb0 < p.num_kb; kb0 += CHUNK_KB, ++chunk) {
                    const uint32_t acc = chunk & 1, aph = (chunk >> 1) & 1;
                    mbar_wait(&acc_empty[acc], aph ^ 1);
                    tc_fence_after();
                    const uint32_t d_big = tmem_base + COL_BIG + acc * BN;
                    const int kb1 = min(kb0 + CHUNK_KB, p.num_kb);
                    for (int kb = kb0; kb < kb1; ++kb, ++it) {
                        const uint32_t os = it % OP_STAGES, oph = (it / OP_STAGES) & 1;
                        mbar_wait(&a_full[os], oph);
                        mbar_wait(&b_full[os], oph);
                        tc_fence_after();
                        const uint32_t bop = bop_ring + os * BOP_STAGE_BYTES;
                        const uint64_t b_hi = make_kmajor_sw128_desc(bop), b_lo = make_kmajor_sw128_desc(bop + TILE_BYTES);
                        const uint32_t a_hi = tmem_base + COL_A + os * 64, a_lo = a_hi + 32;
#pragma unroll
                        for (int kk = 0; kk < BK / 8; ++kk) {
                            const uint64_t adv = static_cast<uint64_t>(kk * 32 >> 4);  // 8 TF32 = 32 bytes along k inside the swizzle row
                            umma_tf32_ts(d_big, a_hi + kk * 8, b_hi + adv, kInstrDesc, (kb == kb0 && kk == 0) ? 0u : 1u);
                            umma_tf32_ts(d_small, a_hi + kk * 8, b_lo + adv, kInstrDesc, (kb == 0 && kk == 0) ? 0u : 1u);
                            umma_tf32_ts(d_small, a_lo + kk * 8, b_hi + adv, kInstrDesc, 1u);
                        }
                        tc_commit(&op_empty[os]);  // A stage in TMEM and B stage in shared memory reusable once these MMAs retire
                    }
                    tc_commit(&acc_full[acc]);
                }
                tc_commit(small_full);
            }
        }
        __syncwarp();
    } else if (wg == 1) {
        // ===== A split: raw tile (shared memory) -> A_hi, A_lo in TMEM; thread = row of the tile =====
        const int r = threadIdx.x - 128;
        const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                const uint32_t rs = it % RAW_STAGES, rph = (it / RAW_STAGES) & 1;
                const uint32_t os = it % OP_STAGES, oph = (it / OP_STAGES) & 1;
                mbar_wait(&raw_full[rs], rph);
                float x[32];
                load_a_row<MODE_A>(raw_ring + rs * RAW_STAGE_BYTES, r, p.conj_a, x);
                uint32_t h[32], l[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) split(x[j], h[j], l[j]);
                mbar_wait(&op_empty[os], oph ^ 1);  // the MMAs that read this TMEM stage have completed
                tc_fence_after();
                const uint32_t dst = tmem_base + lane_sel + COL_A + os * 64;
                tmem_st32(dst, h);
                tmem_st32(dst + 32, l);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&a_full[os]);
                mbar_arrive(&raw_empty[rs]);
            }
        }
    } else if (wg == 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        // ===== B split: raw tile -> B_hi, B_lo operand tiles in shared memory =====
        const int t = threadIdx.x - 256;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                const uint32_t rs = it % RAW_STAGES, rph = (it / RAW_STAGES) & 1;
                const uint32_t os = it % OP_STAGES, oph = (it / OP_STAGES) & 1;
                mbar_wait(&op_empty[os], oph ^ 1);
                mbar_wait(&raw_full[rs], rph);
                const uint32_t bop = bop_ring + os * BOP_STAGE_BYTES;
                split_b<MODE_B>(raw_ring + rs * RAW_STAGE_BYTES + TILE_BYTES, bop, bop + TILE_BYTES, t, p.conj_b);
                fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
                mbar_arrive(&b_full[os]);
                mbar_arrive(&raw_empty[rs]);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        // ===== drain + epilogue: thread = one row of the tile; 128 FP32 running sums in registers =====
        uint32_t chunk = 0, tile_it = 0;
        const bool beta_zero = p.beta[0] == 0.0f && p.beta[1] == 0.0f;
        const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
            int tm, tn;
            tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
            float sum[BN];
#pragma unroll
            for (int j = 0; j < BN; ++j) sum[j] = 0.0f;
            for (int kb0 = 0; kb0 < p.num_kb; kb0 += CHUNK_KB, ++chunk) {
                const uint32_t acc = chunk & 1, aph = (chunk >> 1) & 1;
                mbar_wait(&acc_full[acc], aph);
                tc_fence_after();
                const uint32_t t_big = tmem_base + lane_sel + COL_BIG + acc * BN;
#pragma unroll
                for (int c4 = 0; c4 < BN / 32; ++c4) {
                    uint32_t v[32];
                    tmem_ld32(t_big + c4 * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum[c4 * 32 + j] += __uint_as_float(v[j]);  // round-to-nearest
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[acc]);
            }
            // the tile's cross terms (hi*lo + lo*hi), accumulated over the whole k range in the tensor core
            mbar_wait(small_full, tile_it & 1);
            tc_fence_after();
#pragma unroll
            for (int c4 = 0; c4 < BN / 32; ++c4) {
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_sel + COL_SMALL + c4 * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) sum[c4 * 32 + j] += __uint_as_float(v[j]);
            }
            tc_fence_before();
            mbar_arrive(small_empty);

            const int64_t row = static_cast<int64_t>(tm) * BM + warp * 32 + lane;
            const bool row_ok = row < p.m;
            const int64_t col0 = static_cast<int64_t>(tn) * BN;
#pragma unroll
            for (int j = 0; j < BN; ++j) {
                const int64_t col = col0 + j;
                const bool ok = row_ok && col < p.n;  // the shuffles below run unconditionally (warp-uniform control flow)
                float* c = p.C + row + col * p.ldc;
                const float x = sum[j];
                float out;
                if (!p.cplx) {
                    out = p.alpha[0] * x;
                    if (!beta_zero && ok) out += p.beta[0] * *c;
                } else {
                    // lanes (2i, 2i+1) hold (re, im) of one complex element of C
                    const float other = __shfl_xor_sync(0xffffffffu, x, 1);
                    const bool im = lane & 1;
                    const float xr = im ? other : x, xi = im ? x : other;
                    out = im ? (p.alpha[0] * xi + p.alpha[1] * xr) : (p.alpha[0] * xr - p.alpha[1] * xi);
                    float cown = 0.0f;
                    if (!beta_zero && ok) cown = *c;
                    const float coth = __shfl_xor_sync(0xffffffffu, cown, 1);
                    if (!beta_zero) {
                        const float cr = im ? coth : cown, ci = im ? cown : coth;
                        out += im ? (p.beta[0] * ci + p.beta[1] * cr) : (p.beta[0] * cr - p.beta[1] * ci);
                    }
                }
                if (ok) *c = out;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 13) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

bool make_map_f32(CUtensorMap* map, const float* base, int64_t d0, int64_t d1, int64_t ld, int box0, int box1, bool swizzle128) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box0), static_cast<cuuint32_t>(box1)};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

template <int MODE_A, int MODE_B>
cudaError_t launch(const CUtensorMap& ma, const CUtensorMap& mb, const Params& p, cudaStream_t stream) {
    auto kern = gemm_tf32x3_v2_sm100_kernel<MODE_A, MODE_B>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int tiles = p.tiles_m * p.tiles_n;
    const int grid = gemm_grid(tiles, sm_count());
    kern<<<grid, THREADS, SMEM_BYTES, stream>>>(ma, mb, p);
    return cudaGetLastError();
}

}  // namespace

// The tensor-core path of gemm_tf32x3_sm100.cu::gemm_f32 in its second generation. oa, ob: 0 = N, 1 = T, 2 = C (real types: 0 | 1).
// Returns cudaErrorNotSupported when TMA cannot address the operands (the caller falls back), else the launch status.
cudaError_t gemm_f32_v2_launch(cudaStream_t stream, int elem, int oa, int ob, int64_t m, int64_t n, int64_t k, const float* alpha, const float* A,
                               int64_t lda, const float* B, int64_t ldb, const float* beta, float* C, int64_t ldc) {
    const bool cplx = elem == 2;
    const int64_t e = elem;
    if (get_encode_fn() == nullptr) return cudaErrorNotSupported;
    Params p{};
    p.m = m * e; p.n = n; p.k = k * e;
    p.alpha[0] = alpha[0]; p.alpha[1] = cplx ? alpha[1] : 0.0f;
    p.beta[0] = beta[0]; p.beta[1] = cplx ? beta[1] : 0.0f;
    p.C = C; p.ldc = ldc * e;
    p.tiles_m = static_cast<int>((p.m + BM - 1) / BM);
    p.tiles_n = static_cast<int>((p.n + BN - 1) / BN);
    p.num_kb = static_cast<int>((p.k + BK - 1) / BK);
    p.cplx = cplx ? 1 : 0;
    p.conj_a = oa == 2 ? -1.0f : 1.0f;
    p.conj_b = ob == 2 ? -1.0f : 1.0f;
    CUtensorMap ma, mb;
    bool ok;
    if (cplx) {
        // real views: an interleaved complex matrix with r rows and leading dimension ld is a real matrix with 2r rows and 2 ld
        ok = oa == 0 ? make_map_f32(&ma, A, 2 * m, k, 2 * lda, BM, BK / 2, false) : make_map_f32(&ma, A, 2 * k, m, 2 * lda, BK, BM / 2, true);
        ok = ok && (ob == 0 ? make_map_f32(&mb, B, 2 * k, n, 2 * ldb, BK, BN, true) : make_map_f32(&mb, B, 2 * n, k, 2 * ldb, 2 * BN, BK / 2, false));
        if (!ok) return cudaErrorNotSupported;
        if (oa == 0 && ob == 0) return launch<A_CPLX_ROW, B_KMAJOR>(ma, mb, p, stream);
        if (oa == 0) return launch<A_CPLX_ROW, B_CPLX_ROW>(ma, mb, p, stream);
        if (ob == 0) return launch<A_CPLX_K, B_KMAJOR>(ma, mb, p, stream);
        return launch<A_CPLX_K, B_CPLX_ROW>(ma, mb, p, stream);
    }
    ok = (oa == 0 ? make_map_f32(&ma, A, m, k, lda, BM, BK, false) : make_map_f32(&ma, A, k, m, lda, BK, BM, true)) &&
         (ob == 0 ? make_map_f32(&mb, B, k, n, ldb, BK, BN, true) : make_map_f32(&mb, B, n, k, ldb, BN, BK, false));
    if (!ok) return cudaErrorNotSupported;
    if (oa == 0 && ob == 0) return launch<A_ROWMAJOR, B_KMAJOR>(ma, mb, p, stream);
    if (oa == 0) return launch<A_ROWMAJOR, B_ROWMAJOR>(ma, mb, p, stream);
    if (ob == 0) return launch<A_KMAJOR, B_KMAJOR>(ma, mb, p, stream);
    return launch<A_KMAJOR, B_ROWMAJOR>(ma, mb, p, stream);
}

}  // namespace cosma_b200
