#include <cosma/schedule.hpp>

#include <algorithm>
#include <stdexcept>

namespace cosma {

namespace {
constexpr std::int64_t kAlignElems = 32;  // communication buffers start on 256-byte boundaries (TMA wants >= 16 B)
std::int64_t align_up(std::int64_t v) { return (v + kAlignElems - 1) / kAlignElems * kAlignElems; }
}  // namespace

Schedule::Schedule(const Strategy& strategy, int rank) : strategy_(strategy), rank_(rank) {
    const char labels[3] = {'A', 'B', 'C'};
    const int P = static_cast<int>(strategy_.P);
    for (int x = 0; x < 3; ++x) {
        mappers_[x] = Mapper(labels[x], strategy_, rank < P ? rank : 0);
        MatState& s = st_[x];
        s.bucket_size.assign(P, {});
        s.pointer.assign(P, 0);
        for (int r = 0; r < P; ++r)
            for (const auto& block : mappers_[x].initial_layout(r)) s.bucket_size[r].push_back(static_cast<std::int64_t>(block.size()));
        initial_[x] = rank < P ? static_cast<std::int64_t>(mappers_[x].initial_size(rank)) : 0;
        s.cur_off = 0;
        s.top = align_up(initial_[x]);
        arena_[x] = s.top;
    }
    if (idle() || strategy_.m == 0 || strategy_.n == 0 || strategy_.k == 0) return;
    build(Interval(0, strategy_.m - 1), Interval(0, strategy_.n - 1), Interval(0, strategy_.k - 1), Interval(0, P - 1), 0,
          BetaMode::USER);
}

std::int64_t Schedule::alloc(int x, std::int64_t elements) {
    const std::int64_t off = st_[x].top;
    st_[x].top = align_up(off + elements);
    arena_[x] = std::max(arena_[x], st_[x].top);
    return off;
}

int Schedule::ring_for_step(int step, const Interval& P, int div) {
    for (size_t i = 0; i < rings_.size(); ++i)
        if (rings_[i].step == step) return static_cast<int>(i);
    RingInfo ring;
    ring.step = step;
    int group, offset;
    std::tie(group, offset) = P.locate_in_subinterval(div, rank_);
    ring.my_pos = group;
    ring.color = P.first() + offset;  // rings of one step: disjoint P intervals (multiples of |P|) + offset < |P|/div
    for (int g = 0; g < div; ++g) ring.ranks.push_back(P.first() + P.locate_in_interval(div, g, offset));
    rings_.push_back(ring);
    return static_cast<int>(rings_.size()) - 1;
}

// One node of the recursion (multiply.cpp:317-454): skip the buckets that lie before the current sub-problem,
// then dispatch on the step type.
void Schedule::build(Interval m, Interval n, Interval k, Interval P, size_t step, BetaMode beta) {
    const Interval2D range[3] = {Interval2D(m, k), Interval2D(k, n), Interval2D(m, n)};
    std::vector<int> saved[3];
    std::int64_t shift[3];
    for (int x = 0; x < 3; ++x) {
        MatState& s = st_[x];
        saved[x].assign(s.pointer.begin() + P.first(), s.pointer.begin() + P.last() + 1);
        for (int r = P.first(); r <= P.last(); ++r) {
            const auto& blocks = mappers_[x].initial_layout(r);
            while (s.pointer[r] < static_cast<int>(blocks.size()) && blocks[s.pointer[r]].before(range[x])) ++s.pointer[r];
        }
        shift[x] = 0;
        for (int b = saved[x][rank_ - P.first()]; b < s.pointer[rank_]; ++b) shift[x] += s.bucket_size[rank_][b];
        s.cur_off += shift[x];
    }

    if (strategy_.final_step(step) || strategy_.empty()) {
        ScheduleOp op;
        op.kind = OpKind::GEMM;
        op.a_off = st_[0].cur_off;
        op.b_off = st_[1].cur_off;
        op.c_off = st_[2].cur_off;
        op.m = static_cast<int>(m.length());
        op.n = static_cast<int>(n.length());
        op.k = static_cast<int>(k.length());
        op.beta = beta;
        ops_.push_back(op);
    } else if (strategy_.parallel_step(step)) {
        parallel(m, n, k, P, step, beta);
    } else {
        sequential(m, n, k, P, step, beta);
    }

    for (int x = 0; x < 3; ++x) {
        st_[x].cur_off -= shift[x];
        std::copy(saved[x].begin(), saved[x].end(), st_[x].pointer.begin() + P.first());
    }
}

// multiply.cpp:461-588: all ranks of P solve the sub-problems one after another; for a k split the partial
// products accumulate (beta = 1 from the second sub-problem on).
void Schedule::sequential(Interval m, Interval n, Interval k, Interval P, size_t step, BetaMode beta) {
    const int div = strategy_.divisor(step);
    for (int i = 0; i < div; ++i) {
        if (strategy_.split_m(step)) build(m.subinterval(div, i), n, k, P, step + 1, beta);
        else if (strategy_.split_n(step)) build(m, n.subinterval(div, i), k, P, step + 1, beta);
        else build(m, n, k.subinterval(div, i), P, step + 1, i == 0 ? beta : BetaMode::ONE);
    }
}

// multiply.cpp:648-964
void Schedule::parallel(Interval m, Interval n, Interval k, Interval P, size_t step, BetaMode beta) {
    const int div = strategy_.divisor(step);
    const int part = P.subinterval_index(div, rank_);
    const Interval newP = P.subinterval(div, part);
    const int dm = strategy_.divisor_m(step), dn = strategy_.divisor_n(step), dk = strategy_.divisor_k(step);
    const Interval newm = m.subinterval(dm, dm > 1 ? part : 0);
    const Interval newn = n.subinterval(dn, dn > 1 ? part : 0);
    const Interval newk = k.subinterval(dk, dk > 1 ? part : 0);

    // the matrix that does not contain the split dimension is expanded: n -> A, m -> B, k -> C
    const int x = strategy_.split_n(step) ? 0 : (strategy_.split_m(step) ? 1 : 2);
    const Interval2D range = x == 0 ? Interval2D(m, k) : (x == 1 ? Interval2D(k, n) : Interval2D(m, n));
    MatState& s = st_[x];

    // bucket sizes inside `range` of every rank of P, before the expansion (layout.cpp:160-188)
    const int np = static_cast<int>(P.length()), nnew = static_cast<int>(newP.length());
    std::vector<std::vector<std::int64_t>> before(np);
    for (int r = P.first(); r <= P.last(); ++r) {
        const auto& blocks = mappers_[x].initial_layout(r);
        for (int b = s.pointer[r]; b < static_cast<int>(blocks.size()) && range.contains(blocks[b]); ++b)
            before[r - P.first()].push_back(s.bucket_size[r][b]);
    }
    // after the expansion each bucket of a ring holds the pieces of all its members (layout.cpp:110-132)
    std::vector<std::vector<std::int64_t>> after(nnew);
    for (int ring = 0; ring < nnew; ++ring) {
        after[ring].assign(before[ring].size(), 0);
        for (size_t b = 0; b < before[ring].size(); ++b)
            for (int g = 0; g < div; ++g) after[ring][b] += before[g * nnew + ring][b];
    }
    auto write_sizes = [&](const std::vector<std::vector<std::int64_t>>& sizes, int offset) {
        for (int r = newP.first(); r <= newP.last(); ++r) {
            const auto& v = sizes[r - newP.first() + offset];
            auto& dst = s.bucket_size[r];
            for (size_t i = 0; i < v.size() && s.pointer[r] + i < dst.size(); ++i) dst[s.pointer[r] + i] = v[i];
        }
    };
    write_sizes(after, 0);

    int group, ring_pos;
    std::tie(group, ring_pos) = P.locate_in_subinterval(div, rank_);
    std::int64_t new_size = 0;
    for (auto v : after[ring_pos]) new_size += v;

    const std::int64_t saved_top = s.top;
    const std::int64_t original = s.cur_off;
    const std::int64_t expanded = alloc(x, new_size);
    s.cur_off = expanded;

    ScheduleOp comm;
    comm.matrix = x;
    comm.step = static_cast<int>(step);
    comm.ring_index = ring_for_step(static_cast<int>(step), P, div);
    comm.my_pos = group;
    comm.ring = rings_[comm.ring_index].ranks;
    for (int g = 0; g < div; ++g) comm.piece.push_back(before[g * nnew + ring_pos]);
    comm.regular = comm.piece[0].size() == 1;
    for (int g = 1; g < div && comm.regular; ++g) comm.regular = comm.piece[g] == comm.piece[0];

    BetaMode inner_beta = beta;
    if (x != 2) {
        comm.kind = OpKind::ALLGATHER;
        comm.src_off = original;
        comm.dst_off = expanded;
        ops_.push_back(comm);
    } else if (beta != BetaMode::ZERO) {
        // the reduction happens after the sub-problem, so the sub-problem starts from zero and the caller's beta is
        // applied when the sum comes back (multiply.cpp:794-797). Every expansion gets a fresh arena region here, so
        // the original C is never overwritten by nested rounds and the reference's buffer swap (:868) is not needed.
        inner_beta = BetaMode::ZERO;
    }

    build(newm, newn, newk, newP, step + 1, inner_beta);

    s.cur_off = original;
    if (x == 2) {
        comm.kind = OpKind::REDUCE;
        comm.src_off = expanded;
        comm.dst_off = original;
        comm.beta = beta;
        if (beta != BetaMode::ZERO) {
            std::int64_t mine = 0;
            for (auto v : comm.piece[group]) mine += v;
            comm.tmp_off = alloc(x, mine);
        }
        ops_.push_back(comm);
    }
    s.top = saved_top;  // stack discipline: the expansion (and staging) buffers die with this step
    write_sizes(before, newP.first() - P.first());
}

double Schedule::total_gemm_flops() const {
    double f = 0;
    for (const auto& op : ops_)
        if (op.kind == OpKind::GEMM) f += 2.0 * op.m * op.n * op.k;
    return f;
}

// Flat encoding, one record per op:
//   GEMM      : 0, a_off, b_off, c_off, m, n, k, beta
//   ALLGATHER : 1, matrix, step, ring_index, my_pos, src_off, dst_off, div, nb, regular, ring[div], piece[div][nb]
//   REDUCE    : 2, matrix, step, ring_index, my_pos, src_off, dst_off, tmp_off, beta, div, nb, regular, ring[div], piece[div][nb]
std::vector<std::int64_t> Schedule::serialize() const {
    std::vector<std::int64_t> out;
    for (const auto& op : ops_) {
        out.push_back(static_cast<int>(op.kind));
        if (op.kind == OpKind::GEMM) {
            for (std::int64_t v : {op.a_off, op.b_off, op.c_off, (std::int64_t)op.m, (std::int64_t)op.n, (std::int64_t)op.k,
                                   (std::int64_t)static_cast<int>(op.beta)})
                out.push_back(v);
            continue;
        }
        const std::int64_t div = static_cast<std::int64_t>(op.ring.size());
        const std::int64_t nb = static_cast<std::int64_t>(op.piece[0].size());
        for (std::int64_t v : {(std::int64_t)op.matrix, (std::int64_t)op.step, (std::int64_t)op.ring_index, (std::int64_t)op.my_pos,
                               op.src_off, op.dst_off})
            out.push_back(v);
        if (op.kind == OpKind::REDUCE) {
            out.push_back(op.tmp_off);
            out.push_back(static_cast<int>(op.beta));
        }
        out.push_back(div);
        out.push_back(nb);
        out.push_back(op.regular ? 1 : 0);
        for (int r : op.ring) out.push_back(r);
        for (const auto& member : op.piece)
            for (std::int64_t v : member) out.push_back(v);
    }
    return out;
}

}  // namespace cosma
