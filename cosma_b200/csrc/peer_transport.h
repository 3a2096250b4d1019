// Zero-SM transport for the overlapped schedules (multiply_exec.cu): the ring-of-two transfers of a cosma::OverlapProgram -- the
// allgather pieces and the exchange of the partial C halves -- as COPY-ENGINE peer copies over NVLink into the ring mate's arena
// (CUDA IPC mappings), ordered across the two processes by stream memory operations on epoch flags. No SM is involved, so the GEMM
// panels keep the whole device (no "narrow" launches), and the transfers run at the NVLink copy rate instead of the ~11 GB/s per CTA
// of NCCL kernels squeezed onto a few SMs (profiles/r2_bench_n2_*.json).
//
// Per overlapped communication op a rank owns two flags (device memory, exported with the plan's flag block): ENTERED (written by the
// mate when it has entered multiply call number `epoch`: its previous call has drained, so its landing zones may be overwritten) and
// ARRIVED (written by the mate after its copy into this rank's landing zone has completed). Flags carry the call number and only
// grow; waits are cuStreamWaitValue32(>=) on the communication stream, writes are a local cuStreamWriteValue32 followed by a 4-byte
// copy-engine copy to the mate -- stream-ordered behind the payload copy on the same stream.
#pragma once
#include "exec_internal.h"

namespace cosma_b200 {

bool peer_copy_enabled();  // COSMA_B200_PEER_COPY = ON (default) | OFF
// Collective over the plan's ring communicators (and, for the verdict, over `parent`): exchanges landing zones and flags with every ring
// mate. *ok = false (on every rank alike) when some rank could not set it up; the plan then keeps the NCCL transport.
int peer_transport_setup(Plan& plan, PeerTransport& t, Comm* parent, void* A, void* B, void* C, bool* ok);
void peer_transport_release(PeerTransport& t);
const PeerLink* peer_link(const PeerTransport& t, int micro);

// communication stream `s`: tell the mate of `link` that this rank has entered call `epoch`
int peer_signal_entered(PeerTransport& t, const PeerLink& link, cudaStream_t s);
// wait until the mate has entered, copy `bytes` from src into its landing zone (alt: the beta == 0 zone), then signal ARRIVED
int peer_push(PeerTransport& t, const PeerLink& link, const void* src, size_t bytes, bool alt, cudaStream_t s);
// wait until the mate's copy into this rank's landing zone has completed
int peer_wait_arrived(PeerTransport& t, const PeerLink& link, cudaStream_t s);

}  // namespace cosma_b200
