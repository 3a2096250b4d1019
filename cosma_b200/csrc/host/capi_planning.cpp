// extern "C" access to the host planning layer (Strategy, Mapper): used by the Python mirror, the tests (which pin
// these against the unmodified reference in oracle/_ref) and by foreign-language hosts.
#include "../../../include/cosma_b200.h"
#include <cosma/adapt_strategy.hpp>
#include <cosma/auto_strategy.hpp>
#include <cosma/mapper.hpp>
#include <cosma/strategy.hpp>
#include <costa/grid2grid/comm_volume.hpp>

#include <cstring>
#include <string>

namespace cosma_b200 {
void set_last_error(const std::string& msg);
}

extern "C" {

int cosma_b200_strategy(int m, int n, int k, int P, long long mem_limit, const char* prefix, char* out, int out_len,
                        int* P_out, long long* mem_used) {
    try {
        if (mem_limit <= 0) mem_limit = std::numeric_limits<long long>::max();
        cosma::Strategy st = cosma::parse_strategy(m, n, k, P, prefix ? prefix : "", mem_limit);
        const std::string s = st.to_string();
        if (static_cast<int>(s.size()) + 1 > out_len) return COSMA_B200_INVALID_ARG;
        std::strcpy(out, s.c_str());
        if (P_out) *P_out = static_cast<int>(st.P);
        if (mem_used) *mem_used = st.memory_used;
        return COSMA_B200_OK;
    } catch (const std::exception& e) {
        cosma_b200::set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
}

int cosma_b200_mapper_layout(char label, int m, int n, int k, int P, const char* steps, int* counts, int* out, int out_cap,
                             int* total_blocks) {
    try {
        cosma::Strategy st = cosma::parse_strategy(m, n, k, P, steps ? steps : "");
        cosma::Mapper mapper(label, st, 0);
        int total = 0;
        for (int r = 0; r < mapper.P(); ++r) {
            const auto& blocks = mapper.initial_layout(r);
            counts[r] = static_cast<int>(blocks.size());
            for (const auto& b : blocks) {
                if (4 * (total + 1) > out_cap) return COSMA_B200_INVALID_ARG;
                out[4 * total + 0] = b.rows.first();
                out[4 * total + 1] = b.rows.last();
                out[4 * total + 2] = b.cols.first();
                out[4 * total + 3] = b.cols.last();
                ++total;
            }
        }
        if (total_blocks) *total_blocks = total;
        return COSMA_B200_OK;
    } catch (const std::exception& e) {
        cosma_b200::set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
}

int cosma_b200_mapper_local_coordinates(char label, int m, int n, int k, int P, const char* steps, int gi, int gj,
                                        int64_t* local_idx, int* rank) {
    try {
        cosma::Strategy st = cosma::parse_strategy(m, n, k, P, steps ? steps : "");
        cosma::Mapper mapper(label, st, 0);
        const auto res = mapper.local_coordinates(gi, gj);
        *local_idx = res.first;
        *rank = res.second;
        return COSMA_B200_OK;
    } catch (const std::exception& e) {
        cosma_b200::set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
}

int cosma_b200_mapper_global_coordinates(char label, int m, int n, int k, int P, const char* steps, int64_t local_idx,
                                         int rank, int* gi, int* gj) {
    try {
        cosma::Strategy st = cosma::parse_strategy(m, n, k, P, steps ? steps : "");
        cosma::Mapper mapper(label, st, 0);
        const auto res = mapper.global_coordinates(local_idx, rank);
        *gi = res.first;
        *gj = res.second;
        return COSMA_B200_OK;
    } catch (const std::exception& e) {
        cosma_b200::set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
}

// costa::communication_volume between two grids given as in struct cosma_b200_layout (no local blocks needed) and
// costa::optimal_reordering on a dense volume matrix indexed [min(u,v) * n_ranks + max(u,v)] (see comm_volume.hpp).
int cosma_b200_comm_volume(int rowblocks_a, int colblocks_a, const int* rowsplit_a, const int* colsplit_a, const int* owners_a, int rowblocks_b,
                           int colblocks_b, const int* rowsplit_b, const int* colsplit_b, const int* owners_b, char trans, int n_ranks,
                           long long* volume) {
    try {
        if (!volume || n_ranks < 1) return COSMA_B200_INVALID_ARG;
        auto grid = [n_ranks](int rb, int cb, const int* rs, const int* cs, const int* own) {
            costa::assigned_grid2D g;
            g.grid.rows_split.assign(rs, rs + rb + 1);
            g.grid.cols_split.assign(cs, cs + cb + 1);
            g.owners.assign(own, own + static_cast<size_t>(rb) * cb);
            g.n_ranks = n_ranks;
            return g;
        };
        const costa::comm_volume vol = costa::communication_volume(grid(rowblocks_a, colblocks_a, rowsplit_a, colsplit_a, owners_a),
                                                                   grid(rowblocks_b, colblocks_b, rowsplit_b, colsplit_b, owners_b), trans);
        for (long long i = 0; i < static_cast<long long>(n_ranks) * n_ranks; ++i) volume[i] = 0;
        for (const auto& kv : vol.volume) {
            if (kv.first.src < 0 || kv.first.dest >= n_ranks) return COSMA_B200_INVALID_ARG;
            volume[static_cast<long long>(kv.first.src) * n_ranks + kv.first.dest] = static_cast<long long>(kv.second);
        }
        return COSMA_B200_OK;
    } catch (const std::exception& e) {
        cosma_b200::set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
}

int cosma_b200_optimal_reordering(int n_ranks, const long long* volume, int* permutation, int* reordered) {
    try {
        if (!volume || !permutation || n_ranks < 1) return COSMA_B200_INVALID_ARG;
        costa::comm_volume vol;
        for (int u = 0; u < n_ranks; ++u)
            for (int v = u; v < n_ranks; ++v)
                if (volume[static_cast<long long>(u) * n_ranks + v] > 0)
                    vol.volume[costa::edge_t(u, v)] = static_cast<size_t>(volume[static_cast<long long>(u) * n_ranks + v]);
        bool re = false;
        const std::vector<int> perm = costa::optimal_reordering(vol, n_ranks, re);
        for (int i = 0; i < n_ranks; ++i) permutation[i] = perm[i];
        if (reordered) *reordered = re ? 1 : 0;
        return COSMA_B200_OK;
    } catch (const std::exception& e) {
        cosma_b200::set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
}

// cosma::adapt_strategy_to_block_cyclic_grid for the three 9-int descriptors of a p?gemm call; the prefix ("" = none) is written
// to out. Feed it to cosma_b200_strategy as `prefix` to obtain the completed strategy.
int cosma_b200_adapt_strategy(int m, int n, int k, int P, const int* desca, int ia, int ja, const int* descb, int ib, int jb, const int* descc, int ic,
                              int jc, char transa, char transb, int nprow, int npcol, char order, char* out, int out_len) {
    try {
        if (!desca || !descb || !descc || !out) return COSMA_B200_INVALID_ARG;
        auto of = [](const int* d, int i, int j) {
            cosma::block_cyclic_desc b;
            b.rows = d[2]; b.cols = d[3]; b.block_rows = d[4]; b.block_cols = d[5]; b.i = i; b.j = j;
            return b;
        };
        const std::string s = cosma::adapt_strategy_to_block_cyclic_grid(m, n, k, P, of(desca, ia, ja), of(descb, ib, jb), of(descc, ic, jc), transa, transb,
                                                                         nprow, npcol, order);
        if (static_cast<int>(s.size()) + 1 > out_len) return COSMA_B200_INVALID_ARG;
        std::strcpy(out, s.c_str());
        return COSMA_B200_OK;
    } catch (const std::exception& e) {
        cosma_b200::set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
}

// cosma::fit_strategy_to_memory (auto_strategy.hpp): the strategy whose COMPILED schedule needs at most budget_bytes of device
// memory per rank for elements of elem_bytes; *footprint_bytes = what it needs. INVALID_ARG (with last_error) when nothing fits.
int cosma_b200_fit_strategy(int m, int n, int k, int P, const char* prefix, int elem_bytes, long long budget_bytes, char* out, int out_len,
                            int* P_out, long long* footprint_bytes) {
    try {
        if (!out || elem_bytes < 1) return COSMA_B200_INVALID_ARG;
        const cosma::Strategy st = cosma::fit_strategy_to_memory(m, n, k, static_cast<size_t>(P), prefix ? prefix : "", budget_bytes / elem_bytes);
        const std::string s = st.to_string();
        if (static_cast<int>(s.size()) + 1 > out_len) return COSMA_B200_INVALID_ARG;
        std::strcpy(out, s.c_str());
        if (P_out) *P_out = static_cast<int>(st.P);
        if (footprint_bytes) *footprint_bytes = cosma::schedule_footprint_elements(st) * elem_bytes;
        return COSMA_B200_OK;
    } catch (const std::exception& e) {
        cosma_b200::set_last_error(e.what());
        return COSMA_B200_INVALID_ARG;
    }
}

}  // extern "C"
