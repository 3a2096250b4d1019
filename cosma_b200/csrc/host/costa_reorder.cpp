// Communication-volume graph of a relayout and the rank relabelling that minimises it (see comm_volume.hpp).
#include <costa/grid2grid/comm_volume.hpp>

#include <algorithm>
#include <stdexcept>

namespace costa {

comm_volume communication_volume(const assigned_grid2D& g_init_, const assigned_grid2D& g_final, char trans) {
    const assigned_grid2D g_init = (trans == 'N' || trans == 'n') ? g_init_ : g_init_.transposed();
    if (g_init.num_rows() != g_final.num_rows() || g_init.num_cols() != g_final.num_cols())
        throw std::runtime_error("communication_volume: the grids describe matrices of different shapes");
    // walk both row splits (and both column splits) in lock step: every overlay interval belongs to one block of each grid
    auto overlay = [](const std::vector<int>& a, const std::vector<int>& b) {
        std::vector<int> lines;
        std::merge(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(lines));
        lines.erase(std::unique(lines.begin(), lines.end()), lines.end());
        return lines;
    };
    auto owner_index = [](const std::vector<int>& lines, const std::vector<int>& split) {
        std::vector<int> idx(lines.size() > 0 ? lines.size() - 1 : 0);
        for (size_t c = 0; c + 1 < lines.size(); ++c) idx[c] = static_cast<int>(std::upper_bound(split.begin(), split.end(), lines[c]) - split.begin()) - 1;
        return idx;
    };
    const auto rl = overlay(g_init.grid.rows_split, g_final.grid.rows_split), cl = overlay(g_init.grid.cols_split, g_final.grid.cols_split);
    const auto ri = owner_index(rl, g_init.grid.rows_split), rf = owner_index(rl, g_final.grid.rows_split);
    const auto ci = owner_index(cl, g_init.grid.cols_split), cf = owner_index(cl, g_final.grid.cols_split);
    comm_volume::volume_t w;
    for (size_t r = 0; r + 1 < rl.size(); ++r) {
        const std::size_t h = static_cast<std::size_t>(rl[r + 1] - rl[r]);
        for (size_t c = 0; c + 1 < cl.size(); ++c) {
            const std::size_t area = h * static_cast<std::size_t>(cl[c + 1] - cl[c]);
            if (area == 0) continue;
            w[edge_t(g_init.owner(ri[r], ci[c]), g_final.owner(rf[r], cf[c])).sorted()] += area;
        }
    }
    return comm_volume(std::move(w));
}

std::vector<int> optimal_reordering(const comm_volume& vol, int n_ranks, bool& reordered) {
    std::vector<int> permutation(n_ranks);
    for (int i = 0; i < n_ranks; ++i) permutation[i] = i;
    reordered = false;
    struct cand { long long gain; int u, v; };
    std::vector<cand> edges;
    for (const auto& kv : vol.volume) {
        const int u = kv.first.src, v = kv.first.dest;
        if (u < 0 || v < 0 || u >= n_ranks || v >= n_ranks) throw std::runtime_error("optimal_reordering: rank outside [0, n_ranks)");
        long long gain = static_cast<long long>(kv.second);
        if (u == v) gain = 2 * gain + 1;
        gain -= static_cast<long long>(vol.of(u, u)) + static_cast<long long>(vol.of(v, v));
        if (gain > 0) edges.push_back({gain, u, v});
    }
    std::sort(edges.begin(), edges.end(), [](const cand& a, const cand& b) {
        if (a.gain != b.gain) return a.gain > b.gain;
        return a.u != b.u ? a.u < b.u : a.v < b.v;
    });
    std::vector<char> taken(n_ranks, 0);
    for (const auto& e : edges) {
        if (taken[e.u] || taken[e.v]) continue;
        permutation[e.u] = e.v;
        permutation[e.v] = e.u;
        if (e.u != e.v) reordered = true;
        taken[e.u] = taken[e.v] = 1;
    }
    return permutation;
}

}  // namespace costa
