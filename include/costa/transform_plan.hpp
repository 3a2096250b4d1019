// costa::transform planning -- from a list of (source layout, target layout, op, alpha, beta) build, for ONE rank,
// the three work lists the device executes:
//     pack   : my source pieces that live on another rank  -> contiguous send buffer (plain copy)
//     local  : pieces whose source and target are both mine -> straight into the target block (full transform)
//     unpack : received pieces                              -> target blocks (transpose / conjugate / alpha, beta)
// plus per-peer byte counts of the all-to-all-v in between.
//
// Restates what the reference derives per call in costa::transform (libs/COSTA/src/costa/grid2grid/transform.cpp:
// 231-282): transpose the source grid if op != 'N' (:247-265), overlay the two grids (grid_cover.cpp:54-121), cut every
// local block by the other grid's lines into messages (utils.hpp:26-206), order them by peer (communication_data.cpp:
// 67-164), pack plainly and apply op/alpha/beta when unpacking (communication_data.cpp:166-244). Sender and receiver
// enumerate the overlay cells in the same canonical order, so the packed images agree without exchanging metadata.
#pragma once
#include <costa/erased_layout.hpp>

#include <cstdint>
#include <vector>

namespace costa {

// one rectangular piece, in the argument meaning of the reference's copy_and_transform
// (memory_utils.hpp:287-346): an n_rows x n_cols block OF THE SOURCE, dest = beta*dest + alpha*op(src)
struct piece {
    const void* src = nullptr;    // pack/local: address in a source block; unpack: offset (bytes) into the receive buffer
    void* dst = nullptr;          // unpack/local: address in a target block; pack: offset (bytes) into the send buffer
    std::int64_t src_ld = 0;      // elements; packed pieces are tight (ld = rows if 'C', cols if 'R')
    std::int64_t dst_ld = 0;
    int n_rows = 0, n_cols = 0;   // of the source piece
    char src_ordering = 'C', dst_ordering = 'C';
    bool transpose = false, conjugate = false;
    bool scale_only = false;      // dest = beta*dest, no source (grid_layout::scale_by; beta == 0 stores zeros)
    int transform = 0;            // index into the spec list (selects alpha, beta)
    int peer = 0;                 // pack: destination rank; unpack: source rank; local: me
};

struct transform_spec {
    const erased_layout* from = nullptr;
    const erased_layout* to = nullptr;
    char op = 'N';                // 'N' | 'T' | 'C', applied to the source
    double alpha[2] = {1.0, 0.0};
    double beta[2] = {0.0, 0.0};
};

struct transform_plan {
    int rank = 0, n_ranks = 1, elem_bytes = 8;
    std::vector<piece> pack, local, unpack;
    std::vector<std::int64_t> send_bytes, recv_bytes;  // per peer
    std::vector<std::int64_t> send_off, recv_off;      // per peer, byte offsets into the send / receive buffer
    std::int64_t total_send = 0, total_recv = 0;
    std::vector<transform_spec> specs;                 // alpha/beta/op kept; layout pointers are NOT retained
    std::int64_t local_elements = 0, remote_elements = 0;  // moved by this rank (statistics)
};

// Throws std::runtime_error on inconsistent layouts (dimension mismatch after op, owner outside [0, n_ranks), a block
// that is mine according to the grid but absent from the local block list).
transform_plan plan_transform(const std::vector<transform_spec>& specs, int rank, int n_ranks, int elem_bytes);

}  // namespace costa
