// Glue between the C++ host layer (libcosma.so, plain g++) and the C ABI of the CUDA library (libcosma_b200.so): status ->
// exception, element type -> dtype code, MPI_Comm -> cached NCCL communicator handle, device selection.
#pragma once
#include <cosma/mpi_compat.hpp>
#include <cosma_b200.h>

#include <complex>
#include <stdexcept>
#include <string>

namespace cosma {
namespace b200 {

// throws std::runtime_error("<what>: <cosma_b200_last_error()>") unless status == COSMA_B200_OK
void check(int status, const char* what);
// COSMA_B200_TRACE=ON: one line per host-layer step on stderr ("[cosma rank r] ..."), for locating a stuck collective
bool trace_enabled();
void trace(const char* what);

template <typename T> struct type_code;
template <> struct type_code<float> { static constexpr char value = 's'; };
template <> struct type_code<double> { static constexpr char value = 'd'; };
template <> struct type_code<std::complex<float>> { static constexpr char value = 'c'; };
template <> struct type_code<std::complex<double>> { static constexpr char value = 'z'; };

// (re, im) doubles of a scalar, the convention of the C ABI's plan / layout entry points
template <typename T> inline void to_pair(const T& v, double out[2]) { out[0] = static_cast<double>(v); out[1] = 0.0; }
template <typename T> inline void to_pair(const std::complex<T>& v, double out[2]) { out[0] = v.real(); out[1] = v.imag(); }

// One process drives one GPU: on first use selects device LOCAL_RANK % device_count (the reference does the same per
// rank in its GPU context, libs/Tiled-MM/src/Tiled-MM/mm_handle.cpp). COSMA_B200_KEEP_DEVICE=ON leaves the current device.
void select_device();
// COSMA_B200_BIND_NUMA=ON (applied by select_device): restrict the process to the CPUs local to `device`; false when nothing was changed.
bool bind_to_device_numa_node(int device);

// The NCCL communicator of `comm` (all of its ranks), created collectively on first use -- rank 0 obtains the
// ncclUniqueId and broadcasts it over `comm` as the reference does (src/cosma/gpu/nccl_utils.cpp:21-42) -- and cached by
// communicator identity until release_comm / process exit (the reference caches per context, context.cpp:80-125).
// Call release_comm(comm) before MPI_Comm_free(&comm): communicator identities (MPI_Comm_c2f handles with a real MPI) can be
// handed out again after a free, and a stale cache entry would then be served for the new communicator.
void* comm_handle(MPI_Comm comm);
// The first P ranks of comm as a communicator of their own (comm itself when P == its size), created with
// MPI_Comm_create_group -- collective over those P ranks only, like the reference's communicator (communicator.cpp:282-343)
// -- and cached per (comm, P). Call it only on ranks < P.
MPI_Comm active_comm(MPI_Comm comm, int P);
void release_comm(MPI_Comm comm);
void release_all_comms();

}  // namespace b200
}  // namespace cosma
