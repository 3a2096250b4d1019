// cosma_context / memory_pool (reference src/cosma/context.cpp:11-158, memory_pool.cpp) on top of the C ABI.
#include <cosma/b200_runtime.hpp>
#include <cosma/context.hpp>
#include <cosma/environment_variables.hpp>

#include <complex>
#include <iostream>

namespace cosma {

// ---- memory_pool -----------------------------------------------------------------------------------------------------
template <typename T>
memory_pool<T>::~memory_pool() {
    for (auto& kv : owned_) cosma_b200_host_free(kv.first);
    for (auto& kv : pinned_) cosma_b200_host_unregister(kv.first);
}

template <typename T>
T* memory_pool<T>::allocate(std::size_t n) {
    if (n == 0) return nullptr;
    b200::select_device();
    void* p = nullptr;
    b200::check(cosma_b200_host_alloc(&p, static_cast<uint64_t>(n) * sizeof(T)), "cosma::memory_pool (page-locked host allocation)");
    owned_[static_cast<T*>(p)] = n;
    total_ += n;
    return static_cast<T*>(p);
}

template <typename T>
void memory_pool<T>::deallocate(T* ptr) {
    auto it = owned_.find(ptr);
    if (it == owned_.end()) return;
    total_ -= it->second;
    cosma_b200_host_free(ptr);
    owned_.erase(it);
}

template <typename T>
void memory_pool<T>::pin(T* ptr, std::size_t n) {
    if (!ptr || n == 0 || owned_.count(ptr) || pinned_.count(ptr)) return;
    b200::select_device();
    b200::check(cosma_b200_host_register(ptr, static_cast<uint64_t>(n) * sizeof(T)), "cosma::memory_pool::pin");
    pinned_[ptr] = n;
}

template <typename T>
void memory_pool<T>::unpin_all() {
    for (auto& kv : pinned_) cosma_b200_host_unregister(kv.first);
    pinned_.clear();
}

// ---- cosma_context ---------------------------------------------------------------------------------------------------
template <typename Scalar>
cosma_context<Scalar>::cosma_context() {
    cpu_memory_limit = get_cpu_max_memory<Scalar>();
    adapt_to_scalapack_strategy = get_adapt_strategy();
    overlap_comm_and_comp = get_overlap_comm_and_comp();
}

template <typename Scalar>
cosma_context<Scalar>::cosma_context(size_t cpu_mem_limit, int, int, int, int) : cosma_context() {
    cpu_memory_limit = static_cast<long long>(cpu_mem_limit);
}

template <typename Scalar>
cosma_context<Scalar>::~cosma_context() {
    if (plan_) cosma_b200_plan_destroy(plan_);
}

template <typename Scalar>
void cosma_context<Scalar>::register_state(MPI_Comm comm, const Strategy strategy) {
    if (comm == MPI_COMM_NULL) return;
    const unsigned long long key = comm_key(comm);
    if (plan_ && key == prev_comm_key && strategy == prev_strategy) return;  // context.cpp:89-101: same comm and strategy
    int rank = 0, size = 1;
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &size);
    if (static_cast<int>(strategy.P) > size)
        throw std::runtime_error("cosma: the strategy uses " + std::to_string(strategy.P) + " ranks but the communicator has " + std::to_string(size));
    void* handle = b200::comm_handle(comm);
    if (plan_) {
        b200::check(cosma_b200_stream_synchronize(stream()), "cosma_context (drain before replanning)");
        cosma_b200_plan_destroy(plan_);
        plan_ = nullptr;
    }
    const std::string steps = strategy.to_string();
    b200::trace(("register_state: plan for [" + steps + "]").c_str());
    b200::check(cosma_b200_plan_create_for_strategy(handle, rank, size, strategy.m, strategy.n, strategy.k, static_cast<int>(strategy.P), steps.c_str(),
                                                    b200::type_code<Scalar>::value, &plan_),
                "cosma_context::register_state (plan)");
    b200::trace("register_state: plan ready");
    prev_strategy = strategy;
    prev_comm_key = key;
    if (output && rank == 0) std::cout << "cosma_b200 plan for strategy [" << steps << "] on " << size << " rank(s)" << std::endl;
}

template <typename Scalar>
context<Scalar> make_context() {
    return std::make_unique<cosma_context<Scalar>>();
}
template <typename Scalar>
context<Scalar> make_context(size_t cpu_mem_limit, int streams, int tile_m, int tile_n, int tile_k) {
    return std::make_unique<cosma_context<Scalar>>(cpu_mem_limit, streams, tile_m, tile_n, tile_k);
}
template <typename Scalar>
global_context<Scalar> get_context_instance() {
    // never destroyed: tearing NCCL communicators down during static destruction races with the CUDA runtime's own exit
    static cosma_context<Scalar>* ctxt = new cosma_context<Scalar>();
    return ctxt;
}

#define COSMA_B200_INSTANTIATE(T)                                             \
    template class memory_pool<T>;                                           \
    template class cosma_context<T>;                                         \
    template context<T> make_context<T>();                                   \
    template context<T> make_context<T>(size_t, int, int, int, int);         \
    template global_context<T> get_context_instance<T>();
COSMA_B200_INSTANTIATE(float)
COSMA_B200_INSTANTIATE(double)
COSMA_B200_INSTANTIATE(std::complex<float>)
COSMA_B200_INSTANTIATE(std::complex<double>)
#undef COSMA_B200_INSTANTIATE

}  // namespace cosma
