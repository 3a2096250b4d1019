// cosma::Schedule -- the multiply recursion compiled, once per (strategy, rank), into a flat list of device
// operations over per-matrix arenas.
//
// The reference walks the strategy recursively on EVERY call, shifting host pointers, resizing per-rank bucket
// tables and ping-ponging communication buffers as it goes (src/cosma/multiply.cpp:317-964 with layout.cpp,
// buffer.cpp, matrix.cpp). On a GPU that host work sits between kernels. Here the same recursion is replayed at
// plan time and what remains at run time is a straight-line program of three op kinds,
//     ALLGATHER  (parallel m|n step: expand the un-split operand inside the step's ring)   multiply.cpp:753-786
//     GEMM       (base case, always 'N','N', lda = m, ldb = k, ldc = m)                     multiply.cpp:369-394
//     REDUCE     (parallel k step: reduce-scatter the partial C over the ring, then
//                 C = beta*C + part)                                                          multiply.cpp:909-956
// that the executor issues on CUDA streams (and that tests can interpret on CPU with gloo).
//
// Data placement is the reference's: the initial buffer of each matrix holds this rank's Mapper blocks in order;
// after an expansion the buffer is bucket-major with each bucket the concatenation of the ring members' pieces in
// group order (two_sided_communicator.cpp:104-120), i.e. exactly the column-major sub-matrix. Collectives carry
// exact per-(member, bucket) counts, so neither the reference's padding to the largest piece
// (gpu/nccl_utils.cpp:75-84) nor its post-communication reshuffle copies are needed.
#pragma once
#include <cosma/interval.hpp>
#include <cosma/mapper.hpp>
#include <cosma/strategy.hpp>

#include <cstdint>
#include <vector>

namespace cosma {

enum class OpKind : int { GEMM = 0, ALLGATHER = 1, REDUCE = 2 };
enum class BetaMode : int { ZERO = 0, ONE = 1, USER = 2 };  // beta seen by a GEMM / REDUCE: 0, 1 or the caller's

struct ScheduleOp {
    OpKind kind;
    // GEMM
    std::int64_t a_off = 0, b_off = 0, c_off = 0;  // element offsets into the A / B / C arenas
    int m = 0, n = 0, k = 0;
    BetaMode beta = BetaMode::ZERO;
    // ALLGATHER / REDUCE
    int matrix = 0;            // 0 = A, 1 = B, 2 = C: the matrix being expanded / reduced
    int step = -1;             // strategy step -> which ring communicator
    int ring_index = -1;       // which of this rank's rings (0 .. n_rings-1)
    int my_pos = 0;            // this rank's position (group index) in the ring
    std::int64_t src_off = 0;  // ALLGATHER: this rank's piece; REDUCE: the expanded partial result
    std::int64_t dst_off = 0;  // ALLGATHER: the expanded buffer; REDUCE: where this rank's slice of the sum goes
    std::int64_t tmp_off = -1; // REDUCE with beta != 0: staging for the received sum
    std::vector<int> ring;                             // global ranks of the ring, by group index
    std::vector<std::vector<std::int64_t>> piece;      // piece[member][bucket] element counts
    bool regular = false;      // one bucket and equal pieces: maps to a single ncclAllGather / ncclReduceScatter
};

struct RingInfo {
    int step;                // strategy step
    int color;               // unique among the rings of this step (for a communicator split)
    int my_pos;              // key: group index inside the ring
    std::vector<int> ranks;  // global ranks by group index
};

class Schedule {
  public:
    Schedule() = default;
    // rank >= strategy.P yields an empty schedule (idle rank, multiply.cpp:258-260)
    Schedule(const Strategy& strategy, int rank);

    const Strategy& strategy() const { return strategy_; }
    int rank() const { return rank_; }
    bool idle() const { return rank_ >= static_cast<int>(strategy_.P); }
    const std::vector<ScheduleOp>& ops() const { return ops_; }
    const std::vector<RingInfo>& rings() const { return rings_; }
    // elements each arena must hold (initial buffer first, communication buffers behind it)
    std::int64_t arena_elements(int matrix) const { return arena_[matrix]; }
    std::int64_t initial_elements(int matrix) const { return initial_[matrix]; }
    const Mapper& mapper(int matrix) const { return mappers_[matrix]; }
    double total_gemm_flops() const;  // 2*m*n*k summed over this rank's GEMM ops (x4 for complex by the caller)
    // flat int64 encoding of ops() for foreign-language executors/tests (see schedule.cpp)
    std::vector<std::int64_t> serialize() const;

  private:
    Strategy strategy_;
    int rank_ = 0;
    Mapper mappers_[3];
    std::vector<ScheduleOp> ops_;
    std::vector<RingInfo> rings_;
    std::int64_t arena_[3] = {0, 0, 0};
    std::int64_t initial_[3] = {0, 0, 0};

    // ---- plan-time state (the reference's Layout bookkeeping, for all ranks at once) ----
    struct MatState {
        std::vector<std::vector<std::int64_t>> bucket_size;  // [rank][bucket] current (possibly expanded) sizes
        std::vector<int> pointer;                            // [rank] current bucket
        std::int64_t cur_off = 0;                            // this rank's current matrix position in the arena
        std::int64_t top = 0;                                // arena bump pointer
    };
    MatState st_[3];

    void build(Interval m, Interval n, Interval k, Interval P, size_t step, BetaMode beta);
    void parallel(Interval m, Interval n, Interval k, Interval P, size_t step, BetaMode beta);
    void sequential(Interval m, Interval n, Interval k, Interval P, size_t step, BetaMode beta);
    std::int64_t alloc(int matrix, std::int64_t elements);
    int ring_for_step(int step, const Interval& P, int div);
};

}  // namespace cosma
