#include <cosma/adapt_strategy.hpp>

#include <cctype>

namespace cosma {

std::string adapt_strategy_to_block_cyclic_grid(int m, int n, int k, int P, const block_cyclic_desc& A, const block_cyclic_desc& B,
                                                const block_cyclic_desc& C, char trans_a, char trans_b, int procrows, int proccols, char order) {
    trans_a = static_cast<char>(std::toupper(trans_a));
    trans_b = static_cast<char>(std::toupper(trans_b));
    // the candidate: the largest operand, A before B before C on ties (cosma_pxgemm.cpp:474-499)
    struct candidate {
        long long elements;
        const block_cyclic_desc* d;
        int sub_rows, sub_cols;   // the sub-matrix that takes part in the product, as stored
        const char* dims;         // which problem dimension its rows / columns are
    };
    // NB: for an untransposed B the reference labels rows 'n' and columns 'k' (get_matrix_dimension, :501-515) although B is
    // k x n; the prefix is reproduced as the reference builds it
    const candidate cand[3] = {
        {1LL * m * k, &A, trans_a == 'N' ? m : k, trans_a == 'N' ? k : m, trans_a != 'N' ? "km" : "mk"},
        {1LL * k * n, &B, trans_b == 'N' ? k : n, trans_b == 'N' ? n : k, trans_b != 'N' ? "kn" : "nk"},
        {1LL * m * n, &C, m, n, "mn"},
    };
    int pick = 0;
    for (int x = 1; x < 3; ++x)
        if (cand[x].elements > cand[pick].elements) pick = x;
    const candidate& c = cand[pick];
    if (P < 1 || !(static_cast<double>(c.elements / P) > 1e7)) return "";
    const block_cyclic_desc& d = *c.d;
    if (d.block_rows < 1 || d.block_cols < 1 || procrows < 1 || proccols < 1) return "";
    const bool whole = d.i == 1 && d.j == 1 && c.sub_rows == d.rows && c.sub_cols == d.cols;
    const bool tiled = d.rows % d.block_rows == 0 && d.cols % d.block_cols == 0 && (d.rows / d.block_rows) % procrows == 0 &&
                       (d.cols / d.block_cols) % proccols == 0;
    if (!whole || !tiled) return "";
    std::string prefix;
    auto step = [&prefix](char type, char dim, int div) {
        if (div <= 1) return;
        if (!prefix.empty()) prefix += ',';
        prefix += type;
        prefix += dim;
        prefix += std::to_string(div);
    };
    // how often the process grid repeats along rows / columns: sequential steps
    step('s', c.dims[0], d.rows / d.block_rows / procrows);
    step('s', c.dims[1], d.cols / d.block_cols / proccols);
    // the process grid itself: parallel steps, in the order the ranks are numbered
    if (std::toupper(order) == 'R') {
        step('p', c.dims[0], procrows);
        step('p', c.dims[1], proccols);
    } else {
        step('p', c.dims[1], proccols);
        step('p', c.dims[0], procrows);
    }
    return prefix;
}

}  // namespace cosma
