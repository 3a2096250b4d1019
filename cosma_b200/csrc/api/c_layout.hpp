// erased_layout -> the C ABI's cosma_b200_layout (owning the arrays the C struct points at)
#pragma once
#include <cosma_b200.h>
#include <costa/erased_layout.hpp>

#include <vector>

namespace cosma {
namespace b200 {
struct c_layout {
    std::vector<cosma_b200_block> blocks;
    costa::erased_layout src;  // keeps rows_split / cols_split / owners alive
    cosma_b200_layout c{};
    explicit c_layout(costa::erased_layout&& e) : src(std::move(e)) {
        blocks.reserve(src.blocks.size());
        for (const auto& b : src.blocks) blocks.push_back(cosma_b200_block{b.data, static_cast<int>(b.ld), b.bi, b.bj});
        c.rowblocks = src.grid.grid.n_rows();
        c.colblocks = src.grid.grid.n_cols();
        c.rowsplit = src.grid.grid.rows_split.data();
        c.colsplit = src.grid.grid.cols_split.data();
        c.owners = src.grid.owners.data();
        c.nlocalblocks = static_cast<int>(blocks.size());
        c.localblocks = blocks.data();
    }
    c_layout(const c_layout&) = delete;
    c_layout& operator=(const c_layout&) = delete;
};
}  // namespace b200
}  // namespace cosma
