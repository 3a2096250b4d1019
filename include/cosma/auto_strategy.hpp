// The strategy the library chooses when the caller gives none (p?gemm, multiply_using_layout, cosma_b200_plan_create with
// steps = ""), under the memory limits in effect:
//   COSMA_CPU_MAX_MEMORY  [MB]  the reference's switch (environment_variables.hpp:72-97, context.cpp:11-30): limit of the Strategy's
//                               own memory model, elements per rank -> sequential steps are inserted (strategy.cpp:410-450)
//   COSMA_B200_DEVICE_MEMORY_MB this repo's: the three DEVICE arenas of the COMPILED schedule (local matrices + communication
//                               workspace, what is really allocated in HBM) must fit this many MB per rank; the Strategy limit is
//                               tightened until they do (problems that do not fit 180 GB)
#pragma once
#include <cosma/strategy.hpp>

#include <cstddef>
#include <string>

namespace cosma {

// elements of the A + B + C arenas of the compiled schedule, maximum over the ranks (all of them up to 32 ranks, a sample beyond)
long long schedule_footprint_elements(const Strategy& strategy);

// Strategy(m, n, k, P) completed from `prefix` ("" or leading steps) whose compiled schedule needs at most budget_elements per
// rank; throws std::runtime_error when no sequential splitting makes it fit
Strategy fit_strategy_to_memory(int m, int n, int k, size_t P, const std::string& prefix, long long budget_elements);

// the strategy in effect for (m, n, k, P, steps) and elements of elem_bytes, honouring both switches above
Strategy automatic_strategy(int m, int n, int k, size_t P, const std::string& steps, size_t elem_bytes);

}  // namespace cosma
