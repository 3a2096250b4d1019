#include "host_mirror.h"
#include "exec_internal.h"

#include <algorithm>

namespace cosma_b200 {

bool is_host_pointer(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeUnregistered;
}

HostMirror::~HostMirror() {
    if (slab_) cudaFree(slab_);
}

void HostMirror::add(const void* ptr, size_t pitch, size_t width, size_t height, bool target, bool read_target) {
    if (!ptr || width == 0 || height == 0) return;
    MirrorBlock b;
    b.host = static_cast<char*>(const_cast<void*>(ptr));
    b.pitch = pitch; b.width = width; b.height = height; b.target = target; b.read_target = read_target;
    blocks_.push_back(b);
}

int HostMirror::build() {
    if (built_) return COSMA_B200_OK;
    built_ = true;
    std::vector<MirrorBlock> host;
    for (const auto& b : blocks_)
        if (is_host_pointer(b.host)) host.push_back(b);
    blocks_.swap(host);
    if (blocks_.empty()) return COSMA_B200_OK;
    // merge the byte ranges [host, host + (height-1)*pitch + width)
    std::vector<std::pair<char*, char*>> spans;
    for (const auto& b : blocks_) spans.emplace_back(b.host, b.host + (b.height - 1) * b.pitch + b.width);
    std::sort(spans.begin(), spans.end());
    size_t off = 0;
    for (const auto& s : spans) {
        if (!ranges_.empty() && s.first <= ranges_.back().end) {
            if (s.second > ranges_.back().end) {
                off += static_cast<size_t>(s.second - ranges_.back().end);
                ranges_.back().end = s.second;
            }
            continue;
        }
        off = (off + 255) & ~size_t(255);
        // keep the 16-byte phase of the host address so vectorised requests stay legal on both sides
        off += reinterpret_cast<size_t>(s.first) & 15;
        ranges_.push_back(Range{s.first, s.second, off});
        off += static_cast<size_t>(s.second - s.first);
    }
    slab_bytes_ = off;
    if (cudaMalloc(reinterpret_cast<void**>(&slab_), slab_bytes_) != cudaSuccess) {
        cudaGetLastError();
        set_last_error("host-resident layout: cudaMalloc of the device mirror failed");
        return COSMA_B200_OUT_OF_MEMORY;
    }
    for (auto& b : blocks_) b.dev = static_cast<char*>(translate(b.host));
    return COSMA_B200_OK;
}

void* HostMirror::translate(const void* p) const {
    char* c = static_cast<char*>(const_cast<void*>(p));
    if (!slab_ || ranges_.empty()) return c;
    auto it = std::upper_bound(ranges_.begin(), ranges_.end(), c, [](char* v, const Range& r) { return v < r.begin; });
    if (it == ranges_.begin()) return c;
    --it;
    if (c >= it->end) return c;
    return slab_ + it->dev_off + (c - it->begin);
}

int HostMirror::upload(cudaStream_t stream) const {
    for (const auto& b : blocks_) {
        if (b.target && !b.read_target) continue;
        COSMA_B200_CUDA_TRY(cudaMemcpy2DAsync(b.dev, b.pitch, b.host, b.pitch, b.width, b.height, cudaMemcpyHostToDevice, stream));
    }
    return COSMA_B200_OK;
}

int HostMirror::download(cudaStream_t stream) const {
    for (const auto& b : blocks_) {
        if (!b.target) continue;
        COSMA_B200_CUDA_TRY(cudaMemcpy2DAsync(b.host, b.pitch, b.dev, b.pitch, b.width, b.height, cudaMemcpyDeviceToHost, stream));
    }
    return COSMA_B200_OK;
}

}  // namespace cosma_b200
