// Copy-engine peer transport of the overlapped schedules: see peer_transport.h.
#include "peer_transport.h"

#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace cosma_b200 {

namespace {

constexpr size_t kFlagBlockBytes = 2u << 20;  // an allocation of its own (cudaMalloc sub-allocates smaller requests from shared blocks)
constexpr size_t kScratchOffset = 4096;       // epoch values to copy from: 256 x uint32
constexpr size_t kSendOffset = 65536, kRecvOffset = 131072, kRecBytes = 256;  // staging of the set-up exchange

struct DriverOps {
    CUresult (*wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    CUresult (*write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    CUresult (*address_range)(CUdeviceptr*, size_t*, CUdeviceptr) = nullptr;
    bool ok = false;
};
const DriverOps& driver() {
    static DriverOps d;
    static std::once_flag once;
    std::call_once(once, [] {
        auto get = [](const char* name) -> void* {
            void* sym = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint(name, &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
                (void)cudaGetLastError();
                return nullptr;
            }
            return sym;
        };
        d.wait32 = reinterpret_cast<decltype(d.wait32)>(get("cuStreamWaitValue32"));
        d.write32 = reinterpret_cast<decltype(d.write32)>(get("cuStreamWriteValue32"));
        d.address_range = reinterpret_cast<decltype(d.address_range)>(get("cuMemGetAddressRange"));
        d.ok = d.wait32 && d.write32 && d.address_range;
    });
    return d;
}

// what a rank tells its ring mate about one overlapped communication op
struct Rec {
    cudaIpcMemHandle_t arena;   // the allocation that holds the landing zone(s)
    int64_t off[2];             // byte offsets of the landing zones in it ([1]: the exchange's beta == 0 alternative)
    cudaIpcMemHandle_t flags;   // the sender's flag block
    int64_t flags_off;          // byte offset of this op's flag pair in it
    int32_t ok;
    int32_t pad;
};
static_assert(sizeof(Rec) <= kRecBytes, "staging slot too small");

// process-wide: an IPC handle may be opened once per process; plans share the mappings by reference count
struct Opened {
    void* ptr = nullptr;
    int refs = 0;
};
std::mutex g_ipc_mu;
std::map<std::string, Opened> g_ipc;

void* ipc_open(const cudaIpcMemHandle_t& h, std::string* key_out) {
    const std::string key(reinterpret_cast<const char*>(&h), sizeof(h));
    std::lock_guard<std::mutex> lock(g_ipc_mu);
    auto it = g_ipc.find(key);
    if (it == g_ipc.end()) {
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess || !p) {
            (void)cudaGetLastError();
            return nullptr;
        }
        it = g_ipc.emplace(key, Opened{p, 0}).first;
    }
    ++it->second.refs;
    *key_out = key;
    return it->second.ptr;
}
void ipc_close(const std::string& key) {
    std::lock_guard<std::mutex> lock(g_ipc_mu);
    auto it = g_ipc.find(key);
    if (it == g_ipc.end()) return;
    if (--it->second.refs <= 0) {
        cudaIpcCloseMemHandle(it->second.ptr);
        (void)cudaGetLastError();
        g_ipc.erase(it);
    }
}

// IPC handle of the allocation containing p, and p's byte offset in it
bool export_pointer(const void* p, cudaIpcMemHandle_t* h, int64_t* off) {
    CUdeviceptr base = 0;
    size_t size = 0;
    if (driver().address_range(&base, &size, reinterpret_cast<CUdeviceptr>(p)) != CUDA_SUCCESS) return false;
    if (cudaIpcGetMemHandle(h, reinterpret_cast<void*>(base)) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    *off = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(p) - base);
    return true;
}

uint32_t* scratch_slot(PeerTransport& t) {
    static thread_local unsigned next = 0;
    return reinterpret_cast<uint32_t*>(t.flag_block + kScratchOffset) + (next++ % 256);
}

int write_remote_flag(PeerTransport& t, uint32_t* remote, cudaStream_t s) {
    uint32_t* slot = scratch_slot(t);
    if (driver().write32(reinterpret_cast<CUstream>(s), reinterpret_cast<CUdeviceptr>(slot), t.epoch, 0) != CUDA_SUCCESS) {
        set_last_error("peer transport: cuStreamWriteValue32 failed");
        return COSMA_B200_CUDA_ERROR;
    }
    COSMA_B200_CUDA_TRY(cudaMemcpyAsync(remote, slot, sizeof(uint32_t), cudaMemcpyDefault, s));
    return COSMA_B200_OK;
}

int wait_local_flag(PeerTransport& t, const uint32_t* flag, cudaStream_t s) {
    if (driver().wait32(reinterpret_cast<CUstream>(s), reinterpret_cast<CUdeviceptr>(flag), t.epoch, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) {
        set_last_error("peer transport: cuStreamWaitValue32 failed");
        return COSMA_B200_CUDA_ERROR;
    }
    return COSMA_B200_OK;
}

}  // namespace

bool peer_copy_enabled() {
    const char* v = std::getenv("COSMA_B200_PEER_COPY");
    if (!v || !*v) return true;
    return !(v[0] == 'O' && (v[1] == 'F' || v[1] == 'f')) && !(v[0] == 'o' && v[1] == 'f') && v[0] != '0';
}

void peer_transport_release(PeerTransport& t) {
    for (const auto& key : t.opened) ipc_close(key);
    t.opened.clear();
    t.links.clear();
    if (t.flag_block) cudaFree(t.flag_block);
    t.flag_block = nullptr;
    t.ready = false;
}

const PeerLink* peer_link(const PeerTransport& t, int micro) {
    for (const auto& l : t.links)
        if (l.micro == micro) return &l;
    return nullptr;
}

int peer_transport_setup(Plan& plan, PeerTransport& t, Comm* parent, void* A, void* B, void* C, bool* ok) {
    *ok = false;
    const NcclApi* N = nccl();
    if (!N || !parent || !parent->comm) return COSMA_B200_OK;
    peer_transport_release(t);
    char* arenas[3] = {static_cast<char*>(A), static_cast<char*>(B), static_cast<char*>(C)};
    const int64_t EB = plan.elem_bytes();
    const auto& prog = plan.overlap.ops;
    const auto& ops = plan.schedule.ops();
    bool good = driver().ok;
    cudaStream_t s = nullptr;  // set-up runs on the default stream, synchronously

    // the overlapped communication ops, in program order (ring mates have the same sequence)
    std::vector<int> micro;
    for (size_t i = 0; i < prog.size(); ++i)
        if (prog[i].stream == 1 && (prog[i].kind == cosma::MicroKind::ALLGATHER || prog[i].kind == cosma::MicroKind::EXCHANGE)) micro.push_back(static_cast<int>(i));

    if (!micro.empty()) {
        if (cudaMalloc(reinterpret_cast<void**>(&t.flag_block), kFlagBlockBytes) != cudaSuccess) {
            (void)cudaGetLastError();
            t.flag_block = nullptr;
            set_last_error("peer transport: cannot allocate the flag block");
            return COSMA_B200_OUT_OF_MEMORY;  // nothing to stage the exchange in: a hard error on this rank
        }
        COSMA_B200_CUDA_TRY(cudaMemset(t.flag_block, 0, kFlagBlockBytes));
        if (micro.size() * kRecBytes > kRecvOffset - kSendOffset) good = false;
        // 1. tell every ring mate where its data lands and where this rank's flags are
        std::vector<Rec> mine(micro.size()), theirs(micro.size());
        for (size_t l = 0; l < micro.size(); ++l) {
            const cosma::MicroOp& o = prog[micro[l]];
            Rec& r = mine[l];
            std::memset(&r, 0, sizeof(r));
            const char* zone[2];
            if (o.kind == cosma::MicroKind::ALLGATHER) {
                const auto& op = ops[o.op];
                const int64_t cnt = op.piece[0][0];
                zone[0] = zone[1] = arenas[op.matrix] + (op.dst_off + (1 - op.my_pos) * cnt) * EB;  // the mate's slot of the expanded buffer
            } else {
                zone[0] = arenas[2] + o.recv_off * EB;
                zone[1] = arenas[2] + o.recv_off_zero * EB;
            }
            bool rec_ok = good;
            int64_t off0 = 0, off1 = 0;
            cudaIpcMemHandle_t h1;
            rec_ok = rec_ok && export_pointer(zone[0], &r.arena, &off0) && export_pointer(zone[1], &h1, &off1) &&
                     std::memcmp(&h1, &r.arena, sizeof(h1)) == 0;  // both zones in one allocation (they are parts of the C arena)
            r.off[0] = off0;
            r.off[1] = off1;
            int64_t foff = 0;
            rec_ok = rec_ok && export_pointer(t.flag_block, &r.flags, &foff);
            r.flags_off = foff + static_cast<int64_t>(l) * 2 * sizeof(uint32_t);
            r.ok = rec_ok ? 1 : 0;
            good = good && rec_ok;
        }
        for (size_t l = 0; l < micro.size(); ++l)
            COSMA_B200_CUDA_TRY(cudaMemcpyAsync(t.flag_block + kSendOffset + l * kRecBytes, &mine[l], sizeof(Rec), cudaMemcpyHostToDevice, s));
        for (size_t l = 0; l < micro.size(); ++l) {
            const cosma::MicroOp& o = prog[micro[l]];
            int ring_index, mate;
            if (o.kind == cosma::MicroKind::ALLGATHER) { ring_index = ops[o.op].ring_index; mate = 1 - ops[o.op].my_pos; }
            else { ring_index = o.ring_index; mate = o.peer; }
            ncclComm_t ring = plan.ring_comms[ring_index];
            COSMA_B200_NCCL_TRY(N->GroupStart());
            COSMA_B200_NCCL_TRY(N->Send(t.flag_block + kSendOffset + l * kRecBytes, kRecBytes, ncclChar, mate, ring, s));
            COSMA_B200_NCCL_TRY(N->Recv(t.flag_block + kRecvOffset + l * kRecBytes, kRecBytes, ncclChar, mate, ring, s));
            COSMA_B200_NCCL_TRY(N->GroupEnd());
        }
        COSMA_B200_CUDA_TRY(cudaStreamSynchronize(s));
        for (size_t l = 0; l < micro.size(); ++l)
            COSMA_B200_CUDA_TRY(cudaMemcpy(&theirs[l], t.flag_block + kRecvOffset + l * kRecBytes, sizeof(Rec), cudaMemcpyDeviceToHost));
        // nobody writes a flag before the verdict below (a barrier) has completed on both sides: they are still all zero
        // 2. map the mates' allocations
        for (size_t l = 0; l < micro.size() && good; ++l) {
            const Rec& r = theirs[l];
            if (!r.ok) { good = false; break; }
            std::string ka, kf;
            char* base = static_cast<char*>(ipc_open(r.arena, &ka));
            if (base) t.opened.push_back(ka);
            char* fbase = static_cast<char*>(ipc_open(r.flags, &kf));
            if (fbase) t.opened.push_back(kf);
            if (!base || !fbase) { good = false; break; }
            PeerLink link;
            link.micro = micro[l];
            link.landing[0] = base + r.off[0];
            link.landing[1] = base + r.off[1];
            link.mate_flags = reinterpret_cast<uint32_t*>(fbase + r.flags_off);
            link.my_flags = reinterpret_cast<uint32_t*>(t.flag_block) + 2 * l;
            t.links.push_back(link);
        }
    }
    // 3. one verdict for the whole job (ring mates must use the same transport; idle ranks take part in the reduction only)
    int* d_flag = nullptr;
    COSMA_B200_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_flag), sizeof(int)));
    const int mine_ok = good ? 1 : 0;
    int all_ok = 0;
    cudaMemcpy(d_flag, &mine_ok, sizeof(int), cudaMemcpyHostToDevice);
    ncclResult_t r = N->AllReduce(d_flag, d_flag, 1, ncclInt, ncclMin, parent->comm, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (r == ncclSuccess && e == cudaSuccess) cudaMemcpy(&all_ok, d_flag, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d_flag);
    if (r != ncclSuccess || e != cudaSuccess) {
        peer_transport_release(t);
        set_last_error("peer transport: the verdict reduction failed");
        return COSMA_B200_NCCL_ERROR;
    }
    if (!all_ok) {
        peer_transport_release(t);
        return COSMA_B200_OK;
    }
    t.bound[0] = A; t.bound[1] = B; t.bound[2] = C;
    t.epoch = 0;
    t.ready = true;
    *ok = true;
    return COSMA_B200_OK;
}

int peer_signal_entered(PeerTransport& t, const PeerLink& link, cudaStream_t s) { return write_remote_flag(t, &link.mate_flags[0], s); }

int peer_push(PeerTransport& t, const PeerLink& link, const void* src, size_t bytes, bool alt, cudaStream_t s) {
    int st = wait_local_flag(t, &link.my_flags[0], s);  // the mate has entered this call: its landing zone is free
    if (st != COSMA_B200_OK) return st;
    if (bytes) COSMA_B200_CUDA_TRY(cudaMemcpyAsync(link.landing[alt ? 1 : 0], src, bytes, cudaMemcpyDefault, s));  // copy engine, over NVLink
    return write_remote_flag(t, &link.mate_flags[1], s);
}

int peer_wait_arrived(PeerTransport& t, const PeerLink& link, cudaStream_t s) { return wait_local_flag(t, &link.my_flags[1], s); }

}  // namespace cosma_b200
