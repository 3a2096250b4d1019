// How many CTAs a persistent GEMM launch may use. The GEMM kernels run one CTA per SM; when communication kernels (NCCL) must run
// BESIDE a GEMM (multiply_exec.cu, overlapped schedules), that launch leaves `reserved` SMs free, so that the NCCL kernels find their
// SMs whichever of the two reaches the device first -- and neither waits for the other to end.
#pragma once

namespace cosma_b200 {

int& reserved_sms();  // per host thread; 0 = use every SM (capi.cu)

struct ScopedReservedSms {
    int previous;
    explicit ScopedReservedSms(int r) : previous(reserved_sms()) { reserved_sms() = r; }
    ~ScopedReservedSms() { reserved_sms() = previous; }
};

inline int gemm_grid(int tiles, int sms) {
    int avail = sms - reserved_sms();
    if (avail < 1) avail = 1;
    return tiles < avail ? tiles : avail;
}

}  // namespace cosma_b200
