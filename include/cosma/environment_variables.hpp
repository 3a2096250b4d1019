// Runtime configuration, read from the same environment variables as the reference
// (src/cosma/environment_variables.hpp:10-97). Variables that only steered the reference's host-streaming GPU
// path (COSMA_GPU_STREAMS, COSMA_GPU_MAX_TILE_{M,N,K}, COSMA_GPU_MEMORY_PINNING, COSMA_GPU_UNIFIED_MEMORY,
// COSMA_CPU_MEMORY_ALIGNMENT, COSMA_MEMORY_POOL_AMORTIZATION) are accepted and ignored: operands live in HBM.
#pragma once
#include <limits>
#include <string>

namespace cosma {
bool env_var_defined(const char* name);
bool get_bool_env_var(const std::string& name, bool default_value);   // "ON"/"OFF" (case-insensitive)
int get_int_env_var(const std::string& name, int default_value);
int get_min_local_dimension();        // COSMA_MIN_LOCAL_DIMENSION, default 200
int get_cosma_dim_threshold();        // COSMA_DIM_THRESHOLD, default 0
bool get_adapt_strategy();            // COSMA_ADAPT_STRATEGY, default ON
bool get_overlap_comm_and_comp();     // COSMA_OVERLAP_COMM_AND_COMP, default OFF
// COSMA_CPU_MAX_MEMORY in MB -> number of elements of size elem_bytes; max() if unset. On this build it bounds the
// per-rank DEVICE arena.
long long get_max_memory_elements(std::size_t elem_bytes);
template <typename T>
long long get_cpu_max_memory() { return get_max_memory_elements(sizeof(T)); }
}  // namespace cosma
