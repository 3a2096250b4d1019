// Device mirrors of HOST-resident layout blocks (the reference's calling convention: costa::transform, multiply_using_layout
// and the ScaLAPACK wrappers take host pointers -- libs/COSTA/src/costa/grid2grid/block.hpp:64-137 are views on caller
// memory). A transform plan whose layouts point at host memory gets a device slab that mirrors the byte ranges its blocks
// cover, with the SAME pitch as the host blocks, so translating a piece address is `slab + (addr - range start)` and the
// leading dimensions planned on the host stay valid. Blocks that interleave in one local array (block-cyclic) share a range.
// Per run: H2D of the source blocks (and of the target blocks when some beta != 0), the relayout kernels, D2H of the targets.
#pragma once
#include <cstddef>
#include <vector>
#include <cuda_runtime.h>

namespace cosma_b200 {

struct MirrorBlock {
    char* host = nullptr;
    size_t pitch = 0;   // bytes between consecutive runs (leading dimension)
    size_t width = 0;   // bytes per contiguous run
    size_t height = 0;  // runs
    bool target = false;
    bool read_target = false;  // target that is read (beta != 0): uploaded as well
    char* dev = nullptr;
};

class HostMirror {
  public:
    ~HostMirror();
    void add(const void* ptr, size_t pitch, size_t width, size_t height, bool target, bool read_target);
    // Keeps only the blocks in host memory, merges their byte ranges, allocates the slab. Returns a cosma_b200_status.
    int build();
    bool active() const { return !blocks_.empty() && slab_; }
    // device address mirroring host address p (p itself when it is not inside a mirrored range)
    void* translate(const void* p) const;
    int upload(cudaStream_t stream) const;    // sources + read targets
    int download(cudaStream_t stream) const;  // targets
    size_t slab_bytes() const { return slab_bytes_; }

  private:
    struct Range { char* begin; char* end; size_t dev_off; };
    std::vector<MirrorBlock> blocks_;
    std::vector<Range> ranges_;
    char* slab_ = nullptr;
    size_t slab_bytes_ = 0;
    bool built_ = false;
};

bool is_host_pointer(const void* p);

}  // namespace cosma_b200
