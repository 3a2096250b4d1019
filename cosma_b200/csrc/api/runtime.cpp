#include <cosma/b200_runtime.hpp>
#include <cosma/environment_variables.hpp>

#include <sched.h>

#include <fstream>
#include <sstream>
#include <string>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <utility>
#include <vector>
#include <mutex>

namespace cosma {
namespace b200 {

void check(int status, const char* what) {
    if (status == COSMA_B200_OK) return;
    const char* msg = cosma_b200_last_error();
    throw std::runtime_error(std::string(what) + " failed (status " + std::to_string(status) + ")" + (msg && *msg ? std::string(": ") + msg : ""));
}

bool trace_enabled() {
    static const bool on = get_bool_env_var("COSMA_B200_TRACE", false);
    return on;
}
void trace(const char* what) {
    if (!trace_enabled()) return;
    const char* r = std::getenv("RANK");
    std::fprintf(stderr, "[cosma rank %s] %s\n", r ? r : "0", what);
    std::fflush(stderr);
}

void select_device() {
    static std::once_flag once;
    std::call_once(once, [] {
        if (get_bool_env_var("COSMA_B200_KEEP_DEVICE", false)) return;
        int n = 0;
        check(cosma_b200_device_count(&n), "cosma_b200_device_count");
        if (n < 1) throw std::runtime_error("cosma: no CUDA device visible (there is no CPU fallback)");
        const char* lr = std::getenv("LOCAL_RANK");
        const int local = lr && *lr ? std::atoi(lr) : 0;
        check(cosma_b200_set_device(local % n), "cosma_b200_set_device");
        if (get_bool_env_var("COSMA_B200_BIND_NUMA", false)) bind_to_device_numa_node(local % n);
    });
}

// COSMA_B200_BIND_NUMA=ON: run on the CPUs local to the rank's GPU (sysfs local_cpulist of its PCI function), so that the page-locked
// buffers of the memory pool (first touch) sit next to that GPU's PCIe root -- what `mpirun --bind-to` / numactl does for the
// reference on multi-socket hosts. Best effort: anything missing leaves the affinity as it is.
bool bind_to_device_numa_node(int device) {
    char bdf[32] = {0};
    if (cosma_b200_device_pci_bus_id(device, bdf, sizeof bdf) != COSMA_B200_OK) return false;
    std::ifstream f(std::string("/sys/bus/pci/devices/") + bdf + "/local_cpulist");
    std::string list;
    if (!f || !std::getline(f, list)) return false;
    cpu_set_t allowed, target;
    CPU_ZERO(&allowed);
    CPU_ZERO(&target);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return false;
    int in_target = 0, in_allowed = CPU_COUNT(&allowed);
    std::stringstream parts(list);
    std::string part;
    while (std::getline(parts, part, ',')) {
        if (part.empty()) continue;
        const auto dash = part.find('-');
        const int lo = std::atoi(part.substr(0, dash).c_str());
        const int hi = dash == std::string::npos ? lo : std::atoi(part.substr(dash + 1).c_str());
        for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c)
            if (c >= 0 && CPU_ISSET(c, &allowed) && !CPU_ISSET(c, &target)) { CPU_SET(c, &target); ++in_target; }
    }
    if (in_target == 0 || in_target == in_allowed) return false;
    const bool ok = sched_setaffinity(0, sizeof target, &target) == 0;
    if (ok) trace(("bound to the " + std::to_string(in_target) + " CPUs local to device " + bdf).c_str());
    return ok;
}

namespace {
std::mutex g_mu;
std::map<unsigned long long, void*> g_comms;

void release_key(unsigned long long key) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_comms.find(key);
    if (it == g_comms.end()) return;
    cosma_b200_comm_destroy(it->second);
    g_comms.erase(it);
}

#if defined(COSMA_B200_WITH_MPI)
// MPI recycles communicator handles (and their Fortran indices, the cache key) after MPI_Comm_free / Cblacs_gridexit: a cached NCCL
// communicator must die WITH its MPI communicator, or a later communicator that happens to get the same handle would inherit peers
// that are not its own. An attribute with a delete callback does that -- also for applications that only ever call p?gemm_ and can
// never call release_comm() (the reference needs no such hook: it compares communicators with MPI_Comm_compare, context.cpp:80-125).
int g_keyval = MPI_KEYVAL_INVALID;
int on_comm_free(MPI_Comm, int, void* attribute_val, void*) {
    release_key(static_cast<unsigned long long>(reinterpret_cast<std::uintptr_t>(attribute_val)));
    return MPI_SUCCESS;
}
void attach_to_comm(MPI_Comm comm, unsigned long long key) {
    if (g_keyval == MPI_KEYVAL_INVALID && MPI_Comm_create_keyval(MPI_COMM_NULL_COPY_FN, on_comm_free, &g_keyval, nullptr) != MPI_SUCCESS) return;
    MPI_Comm_set_attr(comm, g_keyval, reinterpret_cast<void*>(static_cast<std::uintptr_t>(key)));
}
#else
void attach_to_comm(MPI_Comm, unsigned long long) {}  // process-group ids are never reused
#endif
}  // namespace

void* comm_handle(MPI_Comm comm) {
    select_device();
    const unsigned long long key = comm_key(comm);
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_comms.find(key);
    if (it != g_comms.end()) return it->second;
    int rank = 0, size = 1;
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &size);
    uint8_t id[128] = {0};
    if (size > 1) {
        trace("comm_handle: ncclGetUniqueId + broadcast");
        if (rank == 0) check(cosma_b200_nccl_unique_id(id), "cosma_b200_nccl_unique_id");
        MPI_Bcast(id, 128, MPI_BYTE, 0, comm);
        trace("comm_handle: ncclCommInitRank");
    }
    void* handle = nullptr;
    check(cosma_b200_comm_create(rank, size, size > 1 ? id : nullptr, &handle), "cosma_b200_comm_create");
    trace("comm_handle: communicator ready");
    g_comms[key] = handle;
    attach_to_comm(comm, key);
    return handle;
}

namespace {
std::map<std::pair<unsigned long long, int>, MPI_Comm> g_active;
}

MPI_Comm active_comm(MPI_Comm comm, int P) {
    int size = 1;
    MPI_Comm_size(comm, &size);
    if (P >= size) return comm;
    const std::pair<unsigned long long, int> key(comm_key(comm), P);
    {
        std::lock_guard<std::mutex> lock(g_mu);
        auto it = g_active.find(key);
        if (it != g_active.end()) return it->second;
    }
    MPI_Group all, first;
    MPI_Comm_group(comm, &all);
    std::vector<int> keep(P);
    for (int i = 0; i < P; ++i) keep[i] = i;
    MPI_Group_incl(all, P, keep.data(), &first);
    MPI_Comm sub = MPI_COMM_NULL;
    MPI_Comm_create_group(comm, first, /*tag=*/P, &sub);
    MPI_Group_free(&all);
    MPI_Group_free(&first);
    std::lock_guard<std::mutex> lock(g_mu);
    g_active[key] = sub;
    return sub;
}

void release_comm(MPI_Comm comm) { release_key(comm_key(comm)); }

void release_all_comms() {
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto& kv : g_comms) cosma_b200_comm_destroy(kv.second);
    g_comms.clear();
}

}  // namespace b200
}  // namespace cosma
