// costa::comm_volume / communication_volume / optimal_reordering (reference libs/COSTA/src/costa/grid2grid/comm_volume.hpp,
// transform.cpp:9-44, ranks_reordering.cpp:4-61): how many ELEMENTS every pair of ranks exchanges when a matrix moves from
// one assigned grid to another, and the rank relabelling that keeps the most data in place.
//
// volume[{u, v}], u <= v, counts the elements that travel between ranks u and v in either direction; {u, u} is what stays
// on rank u. The relabelling is the reference's greedy matching: an edge (u, v) is worth
//     gain(u, v) = vol(u, v) - vol(u, u) - vol(v, v)       (u != v: swapping the labels of u and v makes vol(u, v) local)
//     gain(u, u) = vol(u, u) + 1                           (keeping u in place; the +1 prefers staying on ties)
// (ranks_reordering.cpp:18-31); edges are taken in decreasing gain while both ends are free. Ties are broken by (u, v)
// ascending here, which makes the result independent of hash-map iteration order (the reference's depends on it).
#pragma once
#include <costa/erased_layout.hpp>

#include <cstddef>
#include <map>
#include <utility>
#include <vector>

namespace costa {

struct edge_t {
    int src = 0, dest = 0;
    edge_t() = default;
    edge_t(int s, int d) : src(s), dest(d) {}
    edge_t sorted() const { return src <= dest ? *this : edge_t(dest, src); }
    bool operator==(const edge_t& o) const { return src == o.src && dest == o.dest; }
    bool operator<(const edge_t& o) const { return src != o.src ? src < o.src : dest < o.dest; }
};

struct comm_volume {
    using volume_t = std::map<edge_t, std::size_t>;
    volume_t volume;

    comm_volume() = default;
    explicit comm_volume(volume_t&& v) : volume(std::move(v)) {}
    comm_volume& operator+=(const comm_volume& other) {
        for (const auto& kv : other.volume) volume[kv.first.sorted()] += kv.second;
        return *this;
    }
    comm_volume operator+(const comm_volume& other) const {
        comm_volume r = *this;
        r += other;
        return r;
    }
    std::size_t of(int u, int v) const {
        const auto it = volume.find(edge_t(u, v).sorted());
        return it == volume.end() ? 0 : it->second;
    }
    // elements that cross between different ranks
    std::size_t total_volume() const {
        std::size_t sum = 0;
        for (const auto& kv : volume)
            if (kv.first.src != kv.first.dest) sum += kv.second;
        return sum;
    }
    std::size_t local_volume() const {
        std::size_t sum = 0;
        for (const auto& kv : volume)
            if (kv.first.src == kv.first.dest) sum += kv.second;
        return sum;
    }
};

// trans != 'N': g_init describes the matrix before transposition (its grid is transposed first, transform.cpp:12-13)
comm_volume communication_volume(const assigned_grid2D& g_init, const assigned_grid2D& g_final, char trans);

// permutation[r] = new label of rank r; an involution. reordered = some rank changed its label.
std::vector<int> optimal_reordering(const comm_volume& volume, int n_ranks, bool& reordered);

}  // namespace costa
