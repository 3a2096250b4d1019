// K1/K2: FP64 DGEMM / ZGEMM for sm_100a, device-resident operands, column-major.
//
//   C = alpha * op(A) * op(B) + beta * C        (op = N | T | C)
//
// Replaces the reference's base-case GEMM: local_multiply -> gpu::gemm (Tiled-MM host-streamed
// cublas?gemm), reference src/cosma/local_multiply.cpp:219-269, libs/Tiled-MM/src/Tiled-MM/
// tiled_mm.cpp:181-268,492-624. The reference streams <=5000^3 host tiles over PCIe; here A/B/C
// stay resident in HBM and one persistent kernel does the whole contraction.
//
// Design (see DESIGN.md "K1"):
//  * FP64 on Blackwell is not a tcgen05 kind; the native FP64 tensor instruction is DMMA.8x8x4
//    (mma.sync.m8n8k4.f64), 64 FMA/clk/SM. Accumulators therefore live in registers.
//  * Persistent CTAs (one per SM), 128x128 C tile, BK=16 per pipeline stage, 6 stages (192 KB smem).
//  * Warp-specialised: warp 8 = TMA producer (cp.async.bulk.tensor + mbarrier expect_tx),
//    warps 0-7 = DMMA consumers, each owning a 64(m) x 32(n) register tile (32 DMMA accumulators).
//  * Operand tiles are written by TMA with the 128-byte swizzle, laid out so that every DMMA
//    fragment load (ld.shared.f64, one element per thread) is bank-conflict free; the k index
//    inside each 16-wide k block is permuted per thread (k is a summation index, any bijection
//    that A and B share is legal) to make that possible for both operand layouts at once.
//  * The MMA computes C^T tiles (MMA-M <- n, MMA-N <- m) so that each thread's two accumulator
//    values are adjacent in column-major C and the epilogue uses 16-byte accesses.
//  * alpha/beta in the epilogue; beta == 0 never reads C (ScaLAPACK NaN rule, reference
//    utils/pxgemm_utils.hpp:603-637). transpose/conjugate are folded into the TMA tensor maps and
//    the fragment addressing -- no extra pass, no extra flops.
//  * ZGEMM runs on the same pipeline through the real 2m x 2k x n embedding
//    [Cr;Ci] = [[Ar,-Ai],[Ai,Ar]] * [Br;Bi] evaluated on the fly from interleaved complex tiles:
//    exactly 4 real FMA per complex FMA, i.e. no wasted DMMA work.
#include "sm100_common.cuh"
#include "gemm_f64_sm100.h"
#include "repack.h"
#include "cta_budget.h"

#include <cstdio>
#include <mutex>

namespace cosma_b200 {
namespace {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 16;
constexpr int STAGES = 6;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 4) * 32;  // 2 consumer warpgroups + 1 producer warpgroup
constexpr int WARPS_M = 2;  // warp grid over the CTA tile
constexpr int WARPS_N = 4;
constexpr int WM = BM / WARPS_M;  // 64
constexpr int WN = BN / WARPS_N;  // 32
constexpr int MI = WM / 8;        // 8 fragments along m
constexpr int NI = WN / 8;        // 4 fragments along n
constexpr int OPERAND_STAGE_BYTES = 128 * BK * 8;  // 16 KB per operand per stage
constexpr int STAGE_BYTES = 2 * OPERAND_STAGE_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

// Operand tile layouts in shared memory (written by TMA with CU_TENSOR_MAP_SWIZZLE_128B; "x" is the m or n
// index of the tile, "k" the contraction index, both in REAL (double) units):
//  MN-major ("x is contiguous in global memory"), boxes of 16 x-values by R k-rows (R = 16 real, 8 complex):
//      byte(x,k) = (x/16)*R*128 + k*128 + ((((x%16)>>1) ^ (k&7)) << 4) + (x&1)*8
//  K-major ("k is contiguous in global memory"), one box of X rows by 16 k-values:
//      byte(x,k) = x*128 + (((k>>1) ^ (x&7)) << 4) + (k&1)*8
//
// Complex (ZGEMM) uses the real embedding  [Cr;Ci] = A' * B^  with
//      A'[(i,d)][(l,c)] = +-component(d^c) of op(A)[i][l]   (2m x 2k),   B^[(l,c)][j] = component c of op(B)[l][j]
// so a CTA tile is 128 real rows (64 complex rows of C) x 128 columns and a stage spans 16 real = 8 complex k.
// The A' entries are gathered on the fly from the interleaved (re,im) tile by address arithmetic plus a sign
// flip (integer XOR on the high word): 4 real FMA per complex FMA, no wasted DMMA work.
enum Layout { MN_MAJOR = 0, K_MAJOR = 1 };
enum Op { OP_N = 0, OP_T = 1, OP_C = 2 };

// k permutation: MMA step j (0..3) of a 16-wide k block, thread column t = lane%4 handles
//   k = (t0^j0) | t0<<1 | t1<<2 | (t1^j1)<<3.
// For fixed j the four t's give distinct bits (2,1) [MN-major conflict freedom] and distinct
// bits (3,0) [K-major conflict freedom]; over j = 0..3 every k in 0..15 is used exactly once.
__device__ __forceinline__ int kperm(int t, int j) {
    const int t0 = t & 1, t1 = t >> 1, j0 = j & 1, j1 = j >> 1;
    return (t0 ^ j0) | (t0 << 1) | (t1 << 2) | ((t1 ^ j1) << 3);
}

__device__ __forceinline__ uint32_t off_mn(int x, int k, int rows_per_box) {
    return (x >> 4) * (rows_per_box * 128) + k * 128 + (((((x & 15) >> 1) ^ (k & 7))) << 4) + ((x & 1) << 3);
}
__device__ __forceinline__ uint32_t off_k(int x, int k) {
    return x * 128 + ((((k >> 1) ^ (x & 7))) << 4) + ((k & 1) << 3);
}

// Fragment addressing, split into a per-thread runtime part (depends on lane and k step, hoisted out of all
// loops) and a compile-time part (depends on the fragment index idx):
//      offset(idx, g, kk) = (base(g, kk) ^ (xor64(idx) ? 64 : 0)) + stride(idx)
// A slot: value feeding MMA column m' = idx*8+g at contraction index kk.
template <int OPA, bool CPLX>
__device__ __forceinline__ uint32_t a_frag_base(int g, int kk) {
    if (!CPLX) return OPA == OP_N ? off_mn(g, kk, 16) : off_k(g, kk);
    const int c = kk & 1, d = g & 1;
    if (OPA == OP_N) return off_mn(g ^ c, kk >> 1, 8);  // real row (i, d^c), complex column l
    return off_k(g >> 1, kk ^ d);                       // complex row i, real k index (l, d^c)
}
// ab = stage address + base, abx = ab ^ 64 (stage addresses are 1024-byte aligned, so the xor commutes with the add)
template <int OPA, bool CPLX>
__device__ __forceinline__ uint32_t a_frag_addr(uint32_t ab, uint32_t abx, int idx) {
    if (!CPLX) {
        if (OPA == OP_N) return ((idx & 1) ? abx : ab) + (idx >> 1) * 2048;  // x = idx*8+g, 16-wide boxes
        return ab + idx * 1024;                                              // row x = idx*8+g
    }
    if (OPA == OP_N) return ((idx & 1) ? abx : ab) + (idx >> 1) * 1024;      // boxes of 8 k-rows
    return ((idx & 1) ? abx : ab) + idx * 512;                               // row x = idx*4 + (g>>1)
}
// B slot: value feeding MMA row n = idx*8+g at contraction index kk.
template <int OPB, bool CPLX>
__device__ __forceinline__ uint32_t b_frag_base(int g, int kk) {
    if (!CPLX) return OPB == OP_N ? off_k(g, kk) : off_mn(g, kk, 16);
    if (OPB == OP_N) return off_k(g, kk);               // real k index (l,c), column j
    return off_mn(2 * g + (kk & 1), kk >> 1, 8);        // real row (j,c), complex column l
}
template <int OPB, bool CPLX>
__device__ __forceinline__ uint32_t b_frag_addr(uint32_t bb, uint32_t bbx, int idx) {
    if (!CPLX && OPB != OP_N) return ((idx & 1) ? bbx : bb) + (idx >> 1) * 2048;
    return bb + idx * 1024;  // K-major rows, or complex MN-major (x = idx*16 + 2g + c: one 8-row box per idx)
}

__device__ __forceinline__ double flip_sign(double v, uint32_t mask) {
    return __hiloint2double(__double2hiint(v) ^ static_cast<int>(mask), __double2loint(v));
}

struct GemmParams {
    int64_t m, n, k;      // logical sizes (complex elements for ZGEMM)
    double alpha[2], beta[2];
    double* C;
    int64_t ldc;          // in elements (complex elements for ZGEMM)
    int tiles_m, tiles_n;
    int num_kb;
};

__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int& tm, int& tn) {
    // bands of 16 tile-rows, column-major inside a band: the ~148 tiles in flight at any time
    // cover a ~16 x 9 block, so A row-panels and B column-panels are shared through L2.
    constexpr int G = 16;
    const int band_tiles = G * tiles_n;
    const int band = tile / band_tiles;
    const int rem = tile - band * band_tiles;
    const int rows_in_band = min(G, tiles_m - band * G);
    tn = rem / rows_in_band;
    tm = band * G + rem % rows_in_band;
}

template <int OPA, int OPB, bool CPLX>
__global__ void __launch_bounds__(THREADS, 1)
gemm_f64_sm100_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const GemmParams p) {
    constexpr int LAYOUT_A = (OPA == OP_N) ? MN_MAJOR : K_MAJOR;
    constexpr int LAYOUT_B = (OPB == OP_N) ? K_MAJOR : MN_MAJOR;
    constexpr int A_BYTES = CPLX ? OPERAND_STAGE_BYTES / 2 : OPERAND_STAGE_BYTES;
    constexpr int TX_BYTES = A_BYTES + OPERAND_STAGE_BYTES;

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B needs 1024-byte aligned boxes
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CONSUMER_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const int num_tiles = p.tiles_m * p.tiles_n;
    const int num_kb = p.num_kb;

    if (warp >= CONSUMER_WARPS) {
        // ===== producer warpgroup: give registers back, warp 8 lane 0 drives TMA =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == CONSUMER_WARPS && lane == 0) {
            tma_prefetch_desc(&map_a);
            tma_prefetch_desc(&map_b);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int tm, tn;
                tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
                const int m0 = tm * BM, n0 = tn * BN;  // m0 in real rows (2 per complex row)
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = smem + s * STAGE_BYTES;
                    uint8_t* sb = sa + OPERAND_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[s], TX_BYTES);
                    const int k0 = kb * BK;  // real k units
                    if (!CPLX) {
                        if (LAYOUT_A == MN_MAJOR) {
#pragma unroll
                            for (int b = 0; b < BM / 16; ++b) tma_load_2d(sa + b * 2048, &map_a, &full_bar[s], m0 + 16 * b, k0);
                        } else {
                            tma_load_2d(sa, &map_a, &full_bar[s], k0, m0);
                        }
                        if (LAYOUT_B == MN_MAJOR) {
#pragma unroll
                            for (int b = 0; b < BN / 16; ++b) tma_load_2d(sb + b * 2048, &map_b, &full_bar[s], n0 + 16 * b, k0);
                        } else {
                            tma_load_2d(sb, &map_b, &full_bar[s], k0, n0);
                        }
                    } else {
                        if (LAYOUT_A == MN_MAJOR) {  // real view (2m) x k, boxes {16, 8}
#pragma unroll
                            for (int b = 0; b < BM / 16; ++b) tma_load_2d(sa + b * 1024, &map_a, &full_bar[s], m0 + 16 * b, k0 / 2);
                        } else {                     // real view (2k) x m, box {16, 64}
                            tma_load_2d(sa, &map_a, &full_bar[s], k0, m0 / 2);
                        }
                        if (LAYOUT_B == MN_MAJOR) {  // real view (2n) x k, 16 boxes {16, 8}
#pragma unroll
                            for (int b = 0; b < 2 * BN / 16; ++b)
                                tma_load_2d(sb + b * 1024, &map_b, &full_bar[s], 2 * n0 + 16 * b, k0 / 2);
                        } else {                     // real view (2k) x n, box {16, 128}
                            tma_load_2d(sb, &map_b, &full_bar[s], k0, n0);
                        }
                    }
                }
            }
        }
        return;
    }

    // ===== DMMA consumers =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int g = lane >> 2;
    const int t = lane & 3;
    const int wm = warp % WARPS_M;
    const int wn = warp / WARPS_M;
    const uint32_t smem_base = smem_u32(smem);

    // complex sign masks, indexed by the parity of the k step j (c = (t&1) ^ (j&1))
    uint32_t sign_a[2] = {0u, 0u}, sign_b[2] = {0u, 0u};
    if (CPLX) {
#pragma unroll
        for (int j0 = 0; j0 < 2; ++j0) {
            const int c = (t & 1) ^ j0, d = g & 1;
            // imaginary part is selected when d^c == 1; it enters Cr with -1 (d == 0), Ci with +1; conj flips it
            const bool neg_a = ((d ^ c) == 1) && ((d == 0) != (OPA == OP_C));
            sign_a[j0] = neg_a ? 0x80000000u : 0u;
            sign_b[j0] = (OPB == OP_C && c == 1) ? 0x80000000u : 0u;
        }
    }

    // per-thread fragment address bases for the 4 k steps of a stage (hoisted out of every loop)
    uint32_t a_base[BK / 4], b_base[BK / 4];
#pragma unroll
    for (int j = 0; j < BK / 4; ++j) {
        // lane part + this warp's first fragment (wm*MI and wn*NI are even, so fragment parity is that of mi / ni)
        a_base[j] = a_frag_base<OPA, CPLX>(g, kperm(t, j)) + a_frag_addr<OPA, CPLX>(0u, 0u, wm * MI);
        b_base[j] = OPERAND_STAGE_BYTES + b_frag_base<OPB, CPLX>(g, kperm(t, j)) + b_frag_addr<OPB, CPLX>(0u, 0u, wn * NI);
        // opaque to the optimiser: otherwise ptxas rematerialises the whole swizzle computation (from S2R tid)
        // inside the k loop instead of keeping 8 registers live
        asm volatile("" : "+r"(a_base[j]), "+r"(b_base[j]));
    }

    double acc[NI][MI][2];
    double fa[2][MI], fb[2][NI];
    // loads the fragments of k step j of the stage at shared address `st` into buffer `buf`
    auto load_frags = [&](uint32_t st, int j, int buf) {
        const uint32_t ab = st + a_base[j], abx = ab ^ 64u, bb = st + b_base[j], bbx = bb ^ 64u;
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            fa[buf][mi] = lds_f64(a_frag_addr<OPA, CPLX>(ab, abx, mi));
            if (CPLX) fa[buf][mi] = flip_sign(fa[buf][mi], sign_a[j & 1]);
        }
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            fb[buf][ni] = lds_f64(b_frag_addr<OPB, CPLX>(bb, bbx, ni));
            if (CPLX && OPB == OP_C) fb[buf][ni] = flip_sign(fb[buf][ni], sign_b[j & 1]);
        }
    };

    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int tm, tn;
        tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) acc[ni][mi][0] = acc[ni][mi][1] = 0.0;

        // software pipeline over (stage, k step): the fragments of step j+1 -- crossing into the next stage at
        // j = 3 -- are in flight while the 32 DMMAs of step j issue, so a warp never sits at a stage boundary
        // with an empty DMMA queue.
        mbar_wait(&full_bar[it % STAGES], (it / STAGES) & 1);
        load_frags(smem_base + (it % STAGES) * STAGE_BYTES, 0, 0);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % STAGES;
            const uint32_t st = smem_base + s * STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < BK / 4; ++j) {
                const int cur = j & 1, nxt = cur ^ 1;
                if (j + 1 < BK / 4) {
                    load_frags(st, j + 1, nxt);
                } else if (kb + 1 < num_kb) {
                    const uint32_t itn = it + 1;
                    mbar_wait(&full_bar[itn % STAGES], (itn / STAGES) & 1);
                    load_frags(smem_base + (itn % STAGES) * STAGE_BYTES, 0, nxt);
                }
#pragma unroll
                for (int ni = 0; ni < NI; ++ni)
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi) dmma884(acc[ni][mi][0], acc[ni][mi][1], fb[cur][ni], fa[cur][mi]);
            }
            // every fragment of stage s was consumed by an issued DMMA: hand the slot back to the producer
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }

        // ===== epilogue: C = alpha*acc + beta*C; each thread owns pairs adjacent in column-major C =====
        const int64_t nbase = int64_t(tn) * BN + wn * WN + g;
        if (!CPLX) {
            const double alpha = p.alpha[0], beta = p.beta[0];
            const int64_t mbase = int64_t(tm) * BM + wm * WM + 2 * t;
            const bool vec_ok = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                const int64_t nn = nbase + ni * 8;
                if (nn >= p.n) continue;
                double* ccol = p.C + nn * p.ldc;
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    const int64_t mm = mbase + mi * 8;
                    double v0 = alpha * acc[ni][mi][0];
                    double v1 = alpha * acc[ni][mi][1];
                    if (mm + 1 < p.m && vec_ok) {
                        double2* ptr = reinterpret_cast<double2*>(ccol + mm);
                        if (beta != 0.0) {
                            const double2 old = *ptr;
                            v0 += beta * old.x;
                            v1 += beta * old.y;
                        }
                        *ptr = make_double2(v0, v1);
                    } else {
                        if (mm < p.m) {
                            if (beta != 0.0) v0 += beta * ccol[mm];
                            ccol[mm] = v0;
                        }
                        if (mm + 1 < p.m) {
                            if (beta != 0.0) v1 += beta * ccol[mm + 1];
                            ccol[mm + 1] = v1;
                        }
                    }
                }
            }
        } else {
            const double ar = p.alpha[0], ai = p.alpha[1], br = p.beta[0], bi = p.beta[1];
            const bool beta_zero = (br == 0.0 && bi == 0.0);
            const int64_t ibase = int64_t(tm) * (BM / 2) + wm * (WM / 2) + t;  // complex row
            const bool vec_ok = (reinterpret_cast<uintptr_t>(p.C) & 15) == 0;
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                const int64_t nn = nbase + ni * 8;
                if (nn >= p.n) continue;
                double* ccol = p.C + 2 * nn * p.ldc;
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    const int64_t ii = ibase + mi * 4;
                    if (ii >= p.m) continue;
                    const double xr = acc[ni][mi][0], xi = acc[ni][mi][1];
                    double vr = ar * xr - ai * xi;
                    double vi = ar * xi + ai * xr;
                    double* ptr = ccol + 2 * ii;
                    if (!beta_zero) {
                        double cr, ci;
                        if (vec_ok) { const double2 old = *reinterpret_cast<double2*>(ptr); cr = old.x; ci = old.y; }
                        else { cr = ptr[0]; ci = ptr[1]; }
                        vr += br * cr - bi * ci;
                        vi += br * ci + bi * cr;
                    }
                    if (vec_ok) *reinterpret_cast<double2*>(ptr) = make_double2(vr, vi);
                    else { ptr[0] = vr; ptr[1] = vi; }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Generic path: any leading dimension / alignment (TMA needs 16-byte aligned bases and strides).
// Plain shared-memory tiled DFMA kernel; only odd-lda or unaligned operands come here.
// ---------------------------------------------------------------------------------------------
constexpr int GT = 64, GK = 16;
__global__ void __launch_bounds__(256) dgemm_generic_kernel(int ta, int tb, int64_t m, int64_t n, int64_t k, double alpha,
                                                             const double* __restrict__ A, int64_t lda,
                                                             const double* __restrict__ B, int64_t ldb, double beta,
                                                             double* __restrict__ C, int64_t ldc) {
    __shared__ double sA[GK][GT + 1];
    __shared__ double sB[GK][GT + 1];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int64_t m0 = int64_t(blockIdx.x) * GT, n0 = int64_t(blockIdx.y) * GT;
    double acc[4][4] = {};
    for (int64_t k0 = 0; k0 < k; k0 += GK) {
        for (int e = threadIdx.x; e < GT * GK; e += 256) {
            int x, kk;
            if (ta) { kk = e % GK; x = e / GK; } else { x = e % GT; kk = e / GT; }
            const int64_t gm = m0 + x, gk = k0 + kk;
            double v = 0.0;
            if (gm < m && gk < k) v = ta ? A[gk + gm * lda] : A[gm + gk * lda];
            sA[kk][x] = v;
        }
        for (int e = threadIdx.x; e < GT * GK; e += 256) {
            int x, kk;
            if (tb) { x = e % GT; kk = e / GT; } else { kk = e % GK; x = e / GK; }
            const int64_t gn = n0 + x, gk = k0 + kk;
            double v = 0.0;
            if (gn < n && gk < k) v = tb ? B[gn + gk * ldb] : B[gk + gn * ldb];
            sB[kk][x] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[kk][tx + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[kk][ty + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t gn = n0 + ty + 16 * j;
        if (gn >= n) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t gm = m0 + tx + 16 * i;
            if (gm >= m) continue;
            double v = alpha * acc[i][j];
            if (beta != 0.0) v += beta * C[gm + gn * ldc];
            C[gm + gn * ldc] = v;
        }
    }
}

template <bool CPLX>
__global__ void scale_matrix_kernel(int64_t m, int64_t n, double br, double bi, double* C, int64_t ldc) {
    const int64_t total = m * n;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        if (!CPLX) {
            double* ptr = C + (i % m) + (i / m) * ldc;
            *ptr = (br == 0.0) ? 0.0 : br * *ptr;
        } else {
            double* ptr = C + 2 * ((i % m) + (i / m) * ldc);
            if (br == 0.0 && bi == 0.0) { ptr[0] = 0.0; ptr[1] = 0.0; }
            else { const double cr = ptr[0], ci = ptr[1]; ptr[0] = br * cr - bi * ci; ptr[1] = br * ci + bi * cr; }
        }
    }
}

// complex generic path (only reached with a 16-byte-misaligned base pointer): one thread per C element
__global__ void zgemm_generic_kernel(int oa, int ob, int64_t m, int64_t n, int64_t k, double ar, double ai,
                                     const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                                     double br, double bi, double* __restrict__ C, int64_t ldc) {
    const int64_t i = blockIdx.x * int64_t(16) + threadIdx.x, j = blockIdx.y * int64_t(16) + threadIdx.y;
    if (i >= m || j >= n) return;
    double sr = 0.0, si = 0.0;
    for (int64_t l = 0; l < k; ++l) {
        const double* a = A + 2 * (oa == OP_N ? i + l * lda : l + i * lda);
        const double* b = B + 2 * (ob == OP_N ? l + j * ldb : j + l * ldb);
        const double xr = a[0], xi = (oa == OP_C) ? -a[1] : a[1];
        const double yr = b[0], yi = (ob == OP_C) ? -b[1] : b[1];
        sr += xr * yr - xi * yi;
        si += xr * yi + xi * yr;
    }
    double vr = ar * sr - ai * si, vi = ar * si + ai * sr;
    double* c = C + 2 * (i + j * ldc);
    if (!(br == 0.0 && bi == 0.0)) {
        const double cr = c[0], ci = c[1];
        vr += br * cr - bi * ci;
        vi += br * ci + bi * cr;
    }
    c[0] = vr;
    c[1] = vi;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

// 2-D FP64 tensor map: dim0 (contiguous) extent d0, dim1 extent d1 with stride ld elements.
bool make_map(CUtensorMap* map, const double* base, int64_t d0, int64_t d1, int64_t ld, int box0, int box1) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 8};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box0), static_cast<cuuint32_t>(box1)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int OPA, int OPB, bool CPLX>
cudaError_t launch_tma(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, int sms, cudaStream_t stream) {
    auto kern = gemm_f64_sm100_kernel<OPA, OPB, CPLX>;
    static bool configured = false;  // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int tiles = p.tiles_m * p.tiles_n;
    const int grid = gemm_grid(tiles, sms);
    kern<<<grid, THREADS, SMEM_BYTES, stream>>>(ma, mb, p);
    return cudaGetLastError();
}

template <bool CPLX>
cudaError_t dispatch_tma(int oa, int ob, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, int sms,
                         cudaStream_t st) {
    if (!CPLX) {  // real: 'C' == 'T'
        oa = oa ? OP_T : OP_N;
        ob = ob ? OP_T : OP_N;
    }
#define COSMA_B200_CASE(A_, B_) \
    if (oa == A_ && ob == B_) return launch_tma<A_, B_, CPLX>(ma, mb, p, sms, st);
    COSMA_B200_CASE(OP_N, OP_N) COSMA_B200_CASE(OP_N, OP_T) COSMA_B200_CASE(OP_T, OP_N) COSMA_B200_CASE(OP_T, OP_T)
    if constexpr (CPLX) {
        COSMA_B200_CASE(OP_N, OP_C) COSMA_B200_CASE(OP_T, OP_C) COSMA_B200_CASE(OP_C, OP_N) COSMA_B200_CASE(OP_C, OP_T)
        COSMA_B200_CASE(OP_C, OP_C)
    }
#undef COSMA_B200_CASE
    return cudaErrorInvalidValue;
}

int device_sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

}  // namespace

static int parse_op(char c) {
    switch (c) {
        case 'N': case 'n': return OP_N;
        case 'T': case 't': return OP_T;
        case 'C': case 'c': return OP_C;
        default: return -1;
    }
}

// CPLX == false: double; CPLX == true: complex<double> as interleaved doubles (sizes/lds in complex elements)
template <bool CPLX>
static int gemm_f64_impl(cudaStream_t stream, char transa, char transb, int64_t m, int64_t n, int64_t k,
                         const double* alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                         const double* beta, double* C, int64_t ldc, int* path_used) {
    const int oa = parse_op(transa), ob = parse_op(transb);
    if (oa < 0 || ob < 0 || m < 0 || n < 0 || k < 0) return COSMA_B200_INVALID_ARG;
    const bool ta = oa != OP_N, tb = ob != OP_N;
    auto max1 = [](int64_t v) { return v > 1 ? v : int64_t(1); };
    if (lda < max1(ta ? k : m) || ldb < max1(tb ? n : k) || ldc < max1(m)) return COSMA_B200_INVALID_ARG;
    if (path_used) *path_used = 0;
    if (m == 0 || n == 0) return COSMA_B200_OK;
    constexpr int E = CPLX ? 2 : 1;  // doubles per element
    const bool alpha_zero = alpha[0] == 0.0 && (!CPLX || alpha[1] == 0.0);
    if (k == 0 || alpha_zero) {
        // BLAS semantics: C = beta*C (C not read when beta == 0)
        const bool beta_one = beta[0] == 1.0 && (!CPLX || beta[1] == 0.0);
        if (!beta_one) {
            scale_matrix_kernel<CPLX><<<device_sm_count() * 4, 256, 0, stream>>>(m, n, beta[0], CPLX ? beta[1] : 0.0, C, ldc);
            if (cudaGetLastError() != cudaSuccess) return COSMA_B200_CUDA_ERROR;
        }
        return COSMA_B200_OK;
    }
    const bool aligned = (CPLX || (((lda & 1) == 0) && ((ldb & 1) == 0))) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    if (aligned) {
        CUtensorMap ma, mb;
        bool ok;
        if (!CPLX) {
            // A: 'N' -> stored m x k (m contiguous) MN-major; 'T' -> stored k x m (k contiguous) K-major
            ok = ta ? make_map(&ma, A, k, m, lda, 16, BM) : make_map(&ma, A, m, k, lda, 16, BK);
            // B: 'N' -> stored k x n (k contiguous) K-major; 'T' -> stored n x k (n contiguous) MN-major
            ok = ok && (tb ? make_map(&mb, B, n, k, ldb, 16, BK) : make_map(&mb, B, k, n, ldb, 16, BN));
        } else {
            // real views of the interleaved complex matrices: contiguous extent doubles, leading dimension doubles
            ok = ta ? make_map(&ma, A, 2 * k, m, 2 * lda, 16, BM / 2) : make_map(&ma, A, 2 * m, k, 2 * lda, 16, BK / 2);
            ok = ok && (tb ? make_map(&mb, B, 2 * n, k, 2 * ldb, 16, BK / 2) : make_map(&mb, B, 2 * k, n, 2 * ldb, 16, BN));
        }
        if (ok) {
            GemmParams p;
            p.m = m; p.n = n; p.k = k;
            p.alpha[0] = alpha[0]; p.alpha[1] = CPLX ? alpha[1] : 0.0;
            p.beta[0] = beta[0]; p.beta[1] = CPLX ? beta[1] : 0.0;
            p.C = C; p.ldc = ldc;
            p.tiles_m = static_cast<int>((E * m + BM - 1) / BM);
            p.tiles_n = static_cast<int>((n + BN - 1) / BN);
            p.num_kb = static_cast<int>((E * k + BK - 1) / BK);
            cudaError_t e = dispatch_tma<CPLX>(oa, ob, ma, mb, p, device_sm_count(), stream);
            if (e != cudaSuccess) {
                fprintf(stderr, "cosma_b200: gemm_f64_sm100 launch failed: %s\n", cudaGetErrorString(e));
                return COSMA_B200_CUDA_ERROR;
            }
            if (path_used) *path_used = 1;
            return COSMA_B200_OK;
        }
    }
    // repack what TMA cannot address (repack.h: COSMA_B200_REPACK_UNALIGNED = AUTO | ON | OFF) (odd leading dimension / 8-byte-aligned base) and run the DMMA kernel on the copies
    if (!aligned && repack_wanted(m, n, k)) {
        const bool a_bad = (!CPLX && (lda & 1)) || (reinterpret_cast<uintptr_t>(A) & 15);
        const bool b_bad = (!CPLX && (ldb & 1)) || (reinterpret_cast<uintptr_t>(B) & 15);
        Repacked ra, rb;
        const int eb = CPLX ? 16 : 8;
        cudaError_t e = cudaSuccess;
        if (a_bad) e = repack_operand(stream, A, lda, ta ? k : m, ta ? m : k, eb, 2, ra);
        if (e == cudaSuccess && b_bad) e = repack_operand(stream, B, ldb, tb ? n : k, tb ? k : n, eb, 2, rb);
        if (e == cudaSuccess) {
            const int st = gemm_f64_impl<CPLX>(stream, transa, transb, m, n, k, alpha, a_bad ? static_cast<const double*>(ra.ptr) : A,
                                               a_bad ? ra.ld : lda, b_bad ? static_cast<const double*>(rb.ptr) : B, b_bad ? rb.ld : ldb, beta, C,
                                               ldc, path_used);
            repack_release(stream, ra);
            repack_release(stream, rb);
            return st;
        }
        repack_release(stream, ra);
        repack_release(stream, rb);
        cudaGetLastError();  // scratch not available: fall through to the generic kernel
    }
    if (!CPLX) {
        dim3 grid(static_cast<unsigned>((m + GT - 1) / GT), static_cast<unsigned>((n + GT - 1) / GT));
        dgemm_generic_kernel<<<grid, 256, 0, stream>>>(ta, tb, m, n, k, alpha[0], A, lda, B, ldb, beta[0], C, ldc);
    } else {
        dim3 grid(static_cast<unsigned>((m + 15) / 16), static_cast<unsigned>((n + 15) / 16));
        zgemm_generic_kernel<<<grid, dim3(16, 16), 0, stream>>>(oa, ob, m, n, k, alpha[0], alpha[1], A, lda, B, ldb, beta[0],
                                                                 beta[1], C, ldc);
    }
    if (cudaGetLastError() != cudaSuccess) return COSMA_B200_CUDA_ERROR;
    if (path_used) *path_used = 2;
    return COSMA_B200_OK;
}

int dgemm_sm100(cudaStream_t stream, char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha,
                const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
                int* path_used) {
    return gemm_f64_impl<false>(stream, transa, transb, m, n, k, &alpha, A, lda, B, ldb, &beta, C, ldc, path_used);
}

int zgemm_sm100(cudaStream_t stream, char transa, char transb, int64_t m, int64_t n, int64_t k, const double* alpha,
                const double* A, int64_t lda, const double* B, int64_t ldb, const double* beta, double* C, int64_t ldc,
                int* path_used) {
    return gemm_f64_impl<true>(stream, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, path_used);
}

}  // namespace cosma_b200
