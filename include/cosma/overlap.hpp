// Communication / computation overlap of a compiled cosma::Schedule (COSMA_OVERLAP_COMM_AND_COMP).
//
// The reference overlaps with MPI one-sided transfers issued from a helper thread while the local GEMM works on the
// part of the operands that has already arrived: src/cosma/one_sided_communicator.cpp:417-534 (m split: B arrives
// column block by column block), :540-656 (n split: A arrives k block by k block), :776-1016 (k split: the partial C is
// sent away block by block while the rest is still being computed), gated by strategy.cpp:851-901 and called from
// multiply.cpp:753-786, 909-956.
//
// Here the same idea is a PLAN-TIME lowering of the tail of the op list
//        [ALLGATHER A] [ALLGATHER B] GEMM [REDUCE C]              (rings of two, regular pieces)
// into a short program of micro-ops for two CUDA streams:
//     communication stream : the allgathers at once; later the exchange of the peer's half of the partial C
//     compute stream       : G1 = own k block of A x own column block of B (needs nothing from the network, runs on
//                            SMs - r so that the NCCL kernels always find their r SMs), then the panels that need the
//                            gathered operands, the PEER's half of C first so that it can travel while this rank's own
//                            half is computed, and finally own half += received half.
// Panel widths are chosen in whole waves of the persistent GEMM grid (tile counts that are multiples of the CTA
// count), because a split GEMM pays one partial last wave per launch. The lowering only reorders work along n (independent
// columns) and splits k once (own block first): results are deterministic, equal to the serial schedule on integer-valued
// inputs and within rounding of it otherwise.
#pragma once
#include <cosma/schedule.hpp>

#include <string>
#include <vector>

namespace cosma {

enum class MicroKind : int { GEMM = 0, ALLGATHER = 1, EXCHANGE = 2, ACCUMULATE = 3, SERIAL = 4 };

struct MicroOp {
    MicroKind kind = MicroKind::GEMM;
    int stream = 0;         // 0 = compute, 1 = communication
    std::vector<int> wait;  // micro-ops on the OTHER stream that must have completed first (same-stream order is implicit)
    // GEMM: C(m x n, ldc) = alpha * A(m x k, lda) * B(k x n, ldb) + beta * C, element offsets into the arenas
    std::int64_t a_off = 0, b_off = 0, c_off = 0, lda = 0, ldb = 0, ldc = 0;
    int m = 0, n = 0, k = 0;
    BetaMode beta = BetaMode::ZERO;
    bool narrow = false;  // launch on SMs - reserved_sms: communication kernels run beside this GEMM
    // ALLGATHER / SERIAL: index into Schedule::ops()
    int op = -1;
    // EXCHANGE (ring of two, C arena): send `count` elements from send_off to the peer and receive as many -- at recv_off, or, when
    // `beta` (the reduce's) turns out to be zero at run time, straight at recv_off_zero (the destination: C is never read then)
    int ring_index = -1, peer = 0;
    std::int64_t send_off = 0, recv_off = 0, recv_off_zero = 0, count = 0;
    // ACCUMULATE (C arena): C[dst_off + i] = beta * C[dst_off + i] + C[add_off + i], i < count. beta_term: this is the
    // "beta * C + received" step, skipped when the reduce's beta is zero at run time (the exchange then lands in C itself)
    std::int64_t dst_off = 0, add_off = 0;
    bool beta_term = false;
};

struct OverlapTuning {
    bool enabled = true;
    bool force = false;         // lower whenever the shape of the op list allows it, whatever the estimated gain (tests)
    int sms = 148;              // SMs of the device (the same on every rank)
    int reserved_sms = 8;       // SMs (= NCCL CTAs) left to the communication kernels during narrow GEMMs
    double link_gbps = 84.0;    // NCCL point-to-point rate with `reserved_sms` CTAs (measured: ~10.5 GB/s per CTA)
    double cover = 1.25;        // a narrow GEMM lasts >= cover x the estimated transfer it hides
    bool zero_sm = false;       // the transfers use no SM (copy engines, csrc/peer_transport.h): no narrow launches
    double serial_gbps = 450.0; // what the unconstrained NCCL kernels of the serial schedule reach (for the gain estimate)
    int col_granule = 128;      // panel widths are multiples of this (the GEMM tile width)
    int elem_bytes = 8;
    bool complex_type = false;
    double sm_gflops = 250.0;   // GEMM rate of one SM (FP64 DMMA: 37 TFLOP/s / 148; 3xTF32: ~1080)
};
// COSMA_OVERLAP_COMM_AND_COMP = ON | OFF | FORCE (default ON: lower where the estimate says it pays),
// COSMA_B200_OVERLAP_SMS, COSMA_B200_OVERLAP_GBPS, COSMA_B200_OVERLAP_COVER
OverlapTuning overlap_tuning_from_env(char dtype, int sms);

struct OverlapProgram {
    bool enabled = false;
    std::string why;  // why not, or a one-line description of the lowering
    std::vector<MicroOp> ops;
    double est_serial_ms = 0.0, est_overlap_ms = 0.0, est_comm_ms = 0.0;
    // flat int64 encoding for foreign-language executors / tests: per micro-op
    //   kind, stream, n_wait, wait..., then
    //   GEMM: a_off, b_off, c_off, lda, ldb, ldc, m, n, k, beta, narrow | ALLGATHER, SERIAL: op |
    //   EXCHANGE: ring_index, peer, send_off, recv_off, recv_off_zero, count, beta | ACCUMULATE: dst_off, add_off, count, beta, beta_term
    std::vector<std::int64_t> serialize() const;
};

OverlapProgram plan_overlap(const Schedule& schedule, const OverlapTuning& tuning);

}  // namespace cosma
