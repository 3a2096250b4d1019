#include <costa/transform_plan.hpp>

#include <algorithm>
#include <map>
#include <stdexcept>
#include <string>

namespace costa {

namespace {

constexpr std::int64_t kPieceAlign = 16;     // every packed piece starts on a 16-byte boundary (vector accesses)
constexpr std::int64_t kSegmentAlign = 256;  // every per-peer segment starts on a 256-byte boundary

std::int64_t align_up(std::int64_t v, std::int64_t a) { return (v + a - 1) / a * a; }

// sorted union of two split vectors, duplicates removed (zero-length blocks vanish from the overlay)
std::vector<int> merge_lines(const std::vector<int>& a, const std::vector<int>& b) {
    std::vector<int> out;
    out.reserve(a.size() + b.size());
    std::merge(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(out));
    out.erase(std::unique(out.begin(), out.end()), out.end());
    return out;
}

// for every overlay interval [lines[c], lines[c+1]): the index of the block of `split` that contains it
std::vector<int> covering_block(const std::vector<int>& lines, const std::vector<int>& split) {
    std::vector<int> out(lines.size() > 0 ? lines.size() - 1 : 0);
    for (size_t c = 0; c + 1 < lines.size(); ++c) {
        const auto it = std::upper_bound(split.begin(), split.end(), lines[c]);
        out[c] = static_cast<int>(it - split.begin()) - 1;
    }
    return out;
}

std::map<std::pair<int, int>, const local_block*> index_blocks(const erased_layout& l) {
    std::map<std::pair<int, int>, const local_block*> m;
    for (const auto& b : l.blocks) m[{b.bi, b.bj}] = &b;
    return m;
}

std::int64_t tight_ld(int n_rows, int n_cols, char ordering) { return ordering == 'R' ? n_cols : n_rows; }

}  // namespace

transform_plan plan_transform(const std::vector<transform_spec>& specs, int rank, int n_ranks, int elem_bytes) {
    transform_plan plan;
    plan.rank = rank;
    plan.n_ranks = n_ranks;
    plan.elem_bytes = elem_bytes;
    plan.send_bytes.assign(n_ranks, 0);
    plan.recv_bytes.assign(n_ranks, 0);
    plan.specs = specs;
    for (auto& s : plan.specs) s.from = s.to = nullptr;

    for (size_t si = 0; si < specs.size(); ++si) {
        const transform_spec& spec = specs[si];
        if (!spec.from || !spec.to) throw std::runtime_error("plan_transform: null layout");
        const erased_layout& F = *spec.from;
        const erased_layout& T = *spec.to;
        const char op = spec.op == 'n' ? 'N' : spec.op == 't' ? 'T' : spec.op == 'c' ? 'C' : spec.op;
        if (op != 'N' && op != 'T' && op != 'C') throw std::runtime_error("plan_transform: op must be N, T or C");
        const bool tr = op != 'N';
        const assigned_grid2D FG = tr ? F.grid.transposed() : F.grid;  // the source grid in the target's index space
        const assigned_grid2D& TG = T.grid;
        if (FG.grid.total_rows() != TG.grid.total_rows() || FG.grid.total_cols() != TG.grid.total_cols())
            throw std::runtime_error("plan_transform: op(source) is " + std::to_string(FG.grid.total_rows()) + " x " +
                                     std::to_string(FG.grid.total_cols()) + " but the target is " +
                                     std::to_string(TG.grid.total_rows()) + " x " + std::to_string(TG.grid.total_cols()));
        if (FG.grid.rows_split.front() != 0 || FG.grid.cols_split.front() != 0 || TG.grid.rows_split.front() != 0 ||
            TG.grid.cols_split.front() != 0)
            throw std::runtime_error("plan_transform: split vectors must start at 0");

        const std::vector<int> rows = merge_lines(FG.grid.rows_split, TG.grid.rows_split);
        const std::vector<int> cols = merge_lines(FG.grid.cols_split, TG.grid.cols_split);
        const std::vector<int> f_bi = covering_block(rows, FG.grid.rows_split), t_bi = covering_block(rows, TG.grid.rows_split);
        const std::vector<int> f_bj = covering_block(cols, FG.grid.cols_split), t_bj = covering_block(cols, TG.grid.cols_split);
        const auto f_blocks = index_blocks(F);
        const auto t_blocks = index_blocks(T);

        for (size_t cj = 0; cj + 1 < cols.size(); ++cj) {
            for (size_t ri = 0; ri + 1 < rows.size(); ++ri) {
                const int fo = FG.owner(f_bi[ri], f_bj[cj]);
                const int to = TG.owner(t_bi[ri], t_bj[cj]);
                if (fo < 0 || fo >= n_ranks || to < 0 || to >= n_ranks)
                    throw std::runtime_error("plan_transform: block owner outside [0, n_ranks)");
                if (fo != rank && to != rank) continue;
                const int r0 = rows[ri], r1 = rows[ri + 1], c0 = cols[cj], c1 = cols[cj + 1];
                // the piece in SOURCE coordinates
                const int s_r0 = tr ? c0 : r0, s_c0 = tr ? r0 : c0;
                const int n_rows = tr ? c1 - c0 : r1 - r0, n_cols = tr ? r1 - r0 : c1 - c0;
                const std::int64_t bytes = static_cast<std::int64_t>(n_rows) * n_cols * elem_bytes;

                const void* src_addr = nullptr;
                std::int64_t src_ld = 0;
                if (fo == rank) {
                    const int obi = tr ? f_bj[cj] : f_bi[ri], obj = tr ? f_bi[ri] : f_bj[cj];  // block in the ORIGINAL source grid
                    const auto it = f_blocks.find({obi, obj});
                    if (it == f_blocks.end()) throw std::runtime_error("plan_transform: a source block owned by this rank is missing from its local blocks");
                    const local_block& b = *it->second;
                    const std::int64_t ro = s_r0 - F.grid.grid.rows_split[obi], co = s_c0 - F.grid.grid.cols_split[obj];
                    const std::int64_t off = F.ordering == 'R' ? ro * b.ld + co : ro + co * b.ld;
                    src_addr = static_cast<const char*>(b.data) + off * elem_bytes;
                    src_ld = b.ld;
                }
                void* dst_addr = nullptr;
                std::int64_t dst_ld = 0;
                if (to == rank) {
                    const auto it = t_blocks.find({t_bi[ri], t_bj[cj]});
                    if (it == t_blocks.end()) throw std::runtime_error("plan_transform: a target block owned by this rank is missing from its local blocks");
                    const local_block& b = *it->second;
                    const std::int64_t ro = r0 - TG.grid.rows_split[t_bi[ri]], co = c0 - TG.grid.cols_split[t_bj[cj]];
                    const std::int64_t off = T.ordering == 'R' ? ro * b.ld + co : ro + co * b.ld;
                    dst_addr = static_cast<char*>(b.data) + off * elem_bytes;
                    dst_ld = b.ld;
                }

                piece p;
                p.n_rows = n_rows;
                p.n_cols = n_cols;
                p.transform = static_cast<int>(si);
                if (fo == rank && to == rank) {
                    p.src = src_addr; p.src_ld = src_ld; p.src_ordering = F.ordering;
                    p.dst = dst_addr; p.dst_ld = dst_ld; p.dst_ordering = T.ordering;
                    p.transpose = tr; p.conjugate = op == 'C'; p.peer = rank;
                    plan.local.push_back(p);
                    plan.local_elements += static_cast<std::int64_t>(n_rows) * n_cols;
                } else if (fo == rank) {
                    const std::int64_t at = align_up(plan.send_bytes[to], kPieceAlign);
                    p.src = src_addr; p.src_ld = src_ld; p.src_ordering = F.ordering;
                    p.dst = reinterpret_cast<void*>(at); p.dst_ld = tight_ld(n_rows, n_cols, F.ordering); p.dst_ordering = F.ordering;
                    p.transpose = false; p.conjugate = false; p.transform = -1; p.peer = to;
                    plan.send_bytes[to] = at + bytes;
                    plan.pack.push_back(p);
                    plan.remote_elements += static_cast<std::int64_t>(n_rows) * n_cols;
                } else {
                    const std::int64_t at = align_up(plan.recv_bytes[fo], kPieceAlign);
                    p.src = reinterpret_cast<const void*>(at); p.src_ld = tight_ld(n_rows, n_cols, F.ordering); p.src_ordering = F.ordering;
                    p.dst = dst_addr; p.dst_ld = dst_ld; p.dst_ordering = T.ordering;
                    p.transpose = tr; p.conjugate = op == 'C'; p.peer = fo;
                    plan.recv_bytes[fo] = at + bytes;
                    plan.unpack.push_back(p);
                }
            }
        }
    }

    // segment bases; piece offsets become absolute inside the send / receive buffer
    plan.send_off.assign(n_ranks, 0);
    plan.recv_off.assign(n_ranks, 0);
    std::int64_t s = 0, r = 0;
    for (int p = 0; p < n_ranks; ++p) {
        plan.send_off[p] = s;
        plan.recv_off[p] = r;
        s = align_up(s + plan.send_bytes[p], kSegmentAlign);
        r = align_up(r + plan.recv_bytes[p], kSegmentAlign);
    }
    plan.total_send = s;
    plan.total_recv = r;
    for (auto& p : plan.pack) p.dst = reinterpret_cast<void*>(reinterpret_cast<std::int64_t>(p.dst) + plan.send_off[p.peer]);
    for (auto& p : plan.unpack) p.src = reinterpret_cast<const void*>(reinterpret_cast<std::int64_t>(p.src) + plan.recv_off[p.peer]);
    return plan;
}

}  // namespace costa
