// cosma::cosma_context<Scalar> -- per-type state that outlives a multiply: the memory pool of the matrices and the cached
// (communicator, strategy) -> plan binding. Same public names as the reference (src/cosma/context.hpp:19-97,
// context.cpp:80-158). The plan (compiled schedule + one NCCL ring communicator per parallel step + device arenas)
// plays the role of the reference's `communicator` + Buffer objects; like there, it is rebuilt only when multiply() is
// called with a different communicator or strategy.
#pragma once
#include <cosma/memory_pool.hpp>
#include <cosma/mpi_compat.hpp>
#include <cosma/strategy.hpp>

#include <limits>
#include <memory>

namespace cosma {

template <typename Scalar>
class cosma_context {
  public:
    cosma_context();
    // the tile / stream arguments steered Tiled-MM's host streaming in the reference; operands are HBM-resident here and
    // only cpu_mem_limit (elements, the Strategy memory limit) is kept
    cosma_context(size_t cpu_mem_limit, int streams, int tile_m, int tile_n, int tile_k);
    ~cosma_context();
    cosma_context(const cosma_context&) = delete;
    cosma_context& operator=(const cosma_context&) = delete;

    // (re)binds the context to (comm, strategy): creates the plan unless the cached one matches. Collective over comm.
    void register_state(MPI_Comm comm, const Strategy strategy);

    memory_pool<Scalar>& get_memory_pool() { return memory_pool_; }
    long long get_cpu_memory_limit() const { return cpu_memory_limit; }
    void turn_on_output() { output = true; }

    bool adapt_to_scalapack_strategy = true;
    bool overlap_comm_and_comp = false;
    bool pin_host_buffers = true;

    // B200 side: the current plan handle (cosma_b200_plan_*) and the stream multiply() runs on
    void* plan() const { return plan_; }
    void* stream() const { return nullptr; }
    const Strategy& registered_strategy() const { return prev_strategy; }

  private:
    long long cpu_memory_limit = std::numeric_limits<long long>::max();
    memory_pool<Scalar> memory_pool_;
    bool output = false;
    Strategy prev_strategy;
    unsigned long long prev_comm_key = 0;
    void* plan_ = nullptr;
};

template <typename Scalar>
using global_context = cosma_context<Scalar>*;
template <typename Scalar>
using context = std::unique_ptr<cosma_context<Scalar>>;

template <typename Scalar>
context<Scalar> make_context();
template <typename Scalar>
context<Scalar> make_context(size_t cpu_mem_limit, int streams, int tile_m, int tile_n, int tile_k);
// one lazily created context per Scalar for the whole process (Meyers singleton, as the reference)
template <typename Scalar>
global_context<Scalar> get_context_instance();

}  // namespace cosma
