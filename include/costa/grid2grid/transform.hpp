// costa::transform -- the four overloads of the reference (libs/COSTA/src/costa/grid2grid/transform.hpp:13-43,
// transform.cpp:162-282):  final = beta * final + alpha * op(initial)  for one or several layout pairs in ONE exchange.
// Implemented in libcosma.so over cosma_b200_transform_plan_* (pack kernel -> NCCL all-to-all-v -> unpack kernel).
// Blocks may live in host memory (mirrored through HBM for the call) or in device memory. Collective over comm;
// returns when the result is in place. Throws std::runtime_error on inconsistent layouts.
#pragma once
#include <cosma/mpi_compat.hpp>
#include <costa/grid2grid/grid_layout.hpp>

#include <vector>

namespace costa {

template <typename T>
void transform(grid_layout<T>& initial_layout, grid_layout<T>& final_layout, MPI_Comm comm);

template <typename T>
void transform(grid_layout<T>& initial_layout, grid_layout<T>& final_layout, const char trans, const T alpha, const T beta, MPI_Comm comm);

template <typename T>
void transform(std::vector<layout_ref<T>>& initial_layouts, std::vector<layout_ref<T>>& final_layouts, MPI_Comm comm);

template <typename T>
void transform(std::vector<layout_ref<T>>& initial_layouts, std::vector<layout_ref<T>>& final_layouts, const char* trans, const T* alpha,
               const T* beta, MPI_Comm comm);

}  // namespace costa
