// Internal C++ interface of the batched relayout kernel (R3/R4). The public door is include/cosma_b200.h.
#pragma once
#include <costa/transform.hpp>

#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

namespace cosma_b200 {

// One piece as the kernel sees it: everything normalised to a column-major view.
//   S = source seen column-major, rows x cols with leading dimension src_ld
//   D = S or S^T (flag TRANSPOSE), written column-major with leading dimension dst_ld
struct alignas(16) DevPiece {
    const char* src;
    char* dst;
    std::int64_t src_ld, dst_ld;  // elements
    std::int32_t rows, cols;      // of S
    std::uint32_t flags;
    std::int32_t param;           // index into the scalar table (-1: alpha = 1, beta = 0)
    std::int64_t tile_begin;      // first global tile index of this piece
    std::int64_t pad_;
};
enum : std::uint32_t { PIECE_TRANSPOSE = 1u, PIECE_CONJ = 2u, PIECE_READ_DST = 4u, PIECE_IDENTITY = 8u, PIECE_SCALE_ONLY = 16u };

struct DevScalars {
    double alpha[2], beta[2];
};

constexpr int RELAYOUT_TILE = 32;  // tile edge in elements

struct RelayoutBatch {
    DevPiece* d_pieces = nullptr;      // device
    DevScalars* d_scalars = nullptr;   // device
    int n_pieces = 0;
    std::int64_t total_tiles = 0;
    std::int64_t elements = 0;         // moved per launch (statistics / roofline)
    bool reads_dst = false;
};

// Normalise `pieces` (reference argument meaning, costa::piece) into DevPieces. src_base / dst_base are added to
// piece.src / piece.dst when the piece addresses a buffer by offset (pack: dst, unpack: src).
void relayout_normalise(const std::vector<costa::piece>& pieces, const char* src_base, char* dst_base, int elem_bytes,
                        const std::vector<costa::transform_spec>& specs, std::vector<DevPiece>& out, std::vector<DevScalars>& scalars,
                        std::int64_t* total_tiles, std::int64_t* elements, bool* reads_dst);

// Upload to freshly cudaMalloc'ed device arrays (plan time). Returns a cosma_b200_status.
int relayout_upload(const std::vector<DevPiece>& pieces, const std::vector<DevScalars>& scalars, RelayoutBatch& out);
void relayout_free(RelayoutBatch& b);

// dtype: 's' float, 'd' double, 'c' complex float, 'z' complex double. Asynchronous on `stream`.
int relayout_launch(const RelayoutBatch& batch, char dtype, cudaStream_t stream);

inline int dtype_bytes(char dtype) { return dtype == 's' ? 4 : dtype == 'd' ? 8 : dtype == 'c' ? 8 : dtype == 'z' ? 16 : 0; }

}  // namespace cosma_b200
