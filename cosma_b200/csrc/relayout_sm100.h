// Internal C++ interface of the batched relayout kernel (R3/R4). The public door is include/cosma_b200.h.
#pragma once
#include <costa/transform_plan.hpp>

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
enum : std::uint32_t { PIECE_TRANSPOSE = 1u, PIECE_CONJ = 2u, PIECE_READ_DST = 4u, PIECE_IDENTITY = 8u, PIECE_SCALE_ONLY = 16u, PIECE_VEC16 = 32u };

struct DevScalars {
    double alpha[2], beta[2];
};

// Tile geometry (elements of the column-major source view S), 16 KB per tile:
//   copies     : rows = 1 KB / sizeof(T) (one contiguous kilobyte per column), 16 columns
//   transposes : near-square in BYTES so that the read runs (rows) and the write runs (columns) are both long:
//                16 B: 32 x 32, 8 B: 64 x 32, 4 B: 64 x 64
constexpr int RELAYOUT_TILE_BYTES = 16 * 1024;
__host__ __device__ constexpr int relayout_tile_rows(int elem_bytes, bool transpose) {
    return transpose ? (elem_bytes == 16 ? 32 : 64) : 1024 / elem_bytes;
}
__host__ __device__ constexpr int relayout_tile_cols(int elem_bytes, bool transpose) {
    return RELAYOUT_TILE_BYTES / (relayout_tile_rows(elem_bytes, transpose) * elem_bytes);
}

// Pieces are sorted into four classes, each executed by its own kernel instantiation (one launch per non-empty class):
//   class = (transpose ? 2 : 0) | (16-byte requests legal ? 1 : 0)
constexpr int RELAYOUT_CLASSES = 4;

struct RelayoutHostList {
    std::vector<DevPiece> cls[RELAYOUT_CLASSES];
    std::int64_t tiles[RELAYOUT_CLASSES] = {0, 0, 0, 0};
    std::vector<DevScalars> scalars;
    std::int64_t elements = 0;  // moved per run (statistics / roofline)
    bool reads_dst = false;
};

struct RelayoutBatch {
    DevPiece* d_pieces[RELAYOUT_CLASSES] = {nullptr, nullptr, nullptr, nullptr};  // device
    int n_pieces[RELAYOUT_CLASSES] = {0, 0, 0, 0};
    std::int64_t tiles[RELAYOUT_CLASSES] = {0, 0, 0, 0};
    DevScalars* d_scalars = nullptr;  // device
    std::int64_t elements = 0;
    bool reads_dst = false;
    bool empty() const { return tiles[0] + tiles[1] + tiles[2] + tiles[3] == 0; }
};

// Normalise `pieces` (reference argument meaning, costa::piece) and append them to `out`. src_base / dst_base are added
// to piece.src / piece.dst when the piece addresses a buffer by offset (pack: dst, unpack: src). `specs` supplies
// alpha/beta by piece.transform (set once per list; every call must pass the same specs).
void relayout_normalise(const std::vector<costa::piece>& pieces, const char* src_base, char* dst_base, int elem_bytes,
                        const std::vector<costa::transform_spec>& specs, RelayoutHostList& out);

// Upload to freshly cudaMalloc'ed device arrays (plan time). Returns a cosma_b200_status.
int relayout_upload(const RelayoutHostList& list, RelayoutBatch& out);
void relayout_free(RelayoutBatch& b);

// dtype: 's' float, 'd' double, 'c' complex float, 'z' complex double. Asynchronous on `stream`.
// *launches (optional) += kernels launched.
int relayout_launch(const RelayoutBatch& batch, char dtype, cudaStream_t stream, int* launches = nullptr);

inline int dtype_bytes(char dtype) { return dtype == 's' ? 4 : dtype == 'd' ? 8 : dtype == 'c' ? 8 : dtype == 'z' ? 16 : 0; }

}  // namespace cosma_b200
