// Strategy prefix that mirrors a ScaLAPACK block-cyclic distribution (reference cosma::adapt_strategy_to_block_cyclic_grid,
// src/cosma/cosma_pxgemm.cpp:517-650, called from pxgemm at :162-177 when COSMA_ADAPT_STRATEGY is ON): when the largest of
// op(A), op(B), C is big (more than 1e7 elements per rank) and its block-cyclic grid tiles it perfectly, COSMA is told to
// start with the steps that reproduce that grid -- sequential steps over the block-cycle repetitions, then parallel steps
// over the process rows / columns -- so that this matrix needs (almost) no relayout; the rest of the strategy is then
// completed as usual. Returns the prefix in "-s" notation ("sm8,sk4,pm2,pk4"); empty when the conditions do not hold.
#pragma once
#include <string>

namespace cosma {

struct block_cyclic_desc {
    int rows = 0, cols = 0;          // global matrix size (desc[2], desc[3])
    int block_rows = 0, block_cols = 0;
    int i = 1, j = 1;                // sub-matrix origin, 1-based
};

std::string adapt_strategy_to_block_cyclic_grid(int m, int n, int k, int P, const block_cyclic_desc& A, const block_cyclic_desc& B,
                                                const block_cyclic_desc& C, char trans_a, char trans_b, int procrows, int proccols, char order);

}  // namespace cosma
