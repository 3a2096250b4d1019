#include <cosma/b200_runtime.hpp>
#include <cosma/environment_variables.hpp>

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
    });
}

namespace {
std::mutex g_mu;
std::map<unsigned long long, void*> g_comms;
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

void release_comm(MPI_Comm comm) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_comms.find(comm_key(comm));
    if (it == g_comms.end()) return;
    cosma_b200_comm_destroy(it->second);
    g_comms.erase(it);
}

void release_all_comms() {
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto& kv : g_comms) cosma_b200_comm_destroy(kv.second);
    g_comms.clear();
}

}  // namespace b200
}  // namespace cosma
