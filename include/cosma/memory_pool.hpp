// cosma::memory_pool<T> -- where CosmaMatrix storage comes from. The reference keeps one growing, page-locked host arena
// per context and hands out offsets into it, so matrix_pointer() moves when the pool resizes
// (src/cosma/memory_pool.hpp:25-83, memory_pool.cpp:153-170, multiply.cpp:197-200). Here every buffer is its own
// page-locked allocation (cudaHostAlloc through the C ABI): pointers are stable for the life of the matrix, and the H2D /
// D2H copies of cosma_b200_multiply_host are asynchronous. Communication buffers do not come from this pool -- they are
// device arenas owned by the plan.
#pragma once
#include <cstddef>
#include <unordered_map>

namespace cosma {

template <typename T>
class memory_pool {
  public:
    memory_pool() = default;
    ~memory_pool();
    memory_pool(const memory_pool&) = delete;
    memory_pool& operator=(const memory_pool&) = delete;

    // page-locked buffer of n elements (nullptr for n == 0); throws std::runtime_error when CUDA cannot provide it
    T* allocate(std::size_t n);
    void deallocate(T* ptr);
    // pins / unpins memory the pool does not own (reference memory_pool::pin / unpin_all)
    void pin(T* ptr, std::size_t n);
    void unpin_all();
    std::size_t size() const { return total_; }  // elements currently handed out

  private:
    std::unordered_map<T*, std::size_t> owned_;
    std::unordered_map<T*, std::size_t> pinned_;
    std::size_t total_ = 0;
};

}  // namespace cosma
