// Statistics exchange between the CTAs that share one sample (conv_gn.cuh, conv_in_gn.cu): 8-byte {value, epoch} packets in
// global memory.  A 64-bit relaxed store is single-copy atomic, so a reader that sees the epoch of this launch also sees the
// value: no fence, and no reset between launches (the epoch is a per-launch counter kept next to the packets).
#pragma once
#include "common.cuh"

__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// layout of the caller-provided sync buffer: [0] epoch of the previous launch, [1] CTAs finished, packets from byte 64 on
#define XCHG_MAX_G 32
__device__ __forceinline__ unsigned long long* xchg_slot(void* sync, int b, int src) {
    return reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(sync) + 64) + ((size_t)b * XCHG_MAX_G + src) * 16;
}
// last CTA out advances the epoch for the next launch (every CTA read it before it could finish)
__device__ __forceinline__ void xchg_finish(unsigned int* ctrl) {
    __threadfence();
    const unsigned int done = atomicAdd(ctrl + 1, 1u);
    if (done == gridDim.x - 1) {
        ctrl[1] = 0u;
        __threadfence();
        atomicAdd(ctrl, 1u);
    }
}
