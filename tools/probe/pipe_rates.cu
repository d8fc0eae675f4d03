// Probe: per-SMSP issue cost (cycles per warp instruction) of the instructions the fused epilogue is made of.
// 1 CTA, W warps (W/4 per scheduler), each warp runs N independent copies of the op in a loop; cycles = clock64 delta.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#define ITERS 512
#define UNR 8

template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
    __shared__ __align__(16) float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed + i;
    __syncthreads();
    float x[UNR];
    unsigned long long p[UNR];
    uint32_t h[UNR];
    for (int i = 0; i < UNR; ++i) { x[i] = seed + threadIdx.x * 0.001f + i; p[i] = (unsigned long long)__float_as_uint(x[i]) << 32 | __float_as_uint(x[i]); h[i] = 0x3c003c00u + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNR; ++i) {
            if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 2) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
            if (OP == 3) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h[i]));
            if (OP == 4) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[i]));
            if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
            if (OP == 6) asm volatile("shfl.sync.bfly.b32 %0, %0, 4, 0x1f, 0xffffffff;" : "+r"(h[i]));
            if (OP == 7) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(x[i]), "f"(x[(i + 1) % UNR]));
            if (OP == 8) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(sm) + (uint32_t)(i * 16))); x[i] += v.x; }
            if (OP == 9) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(sm) + (uint32_t)(i * 64 + (threadIdx.x & 3) * 16))); x[i] += v.x; }
            if (OP == 10) asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %1, %1, %1};" ::"r"((uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 128 + (i & 7) * 16), "r"(h[i]) : "memory");
            if (OP == 11) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 12) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(x[i]), "f"(x[(i + 1) % UNR]));
        }
    }
    const long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < UNR; ++i) acc += x[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int warps) {
    float* out; long long* cyc;
    cudaMalloc(&out, 4 * 1024 * 148); cudaMalloc(&cyc, 8 * 148);
    k<OP><<<1, warps * 32>>>(out, cyc, 0.5f);
    k<OP><<<1, warps * 32>>>(out, cyc, 0.5f);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_sched = (double)c / ((double)ITERS * UNR * (warps / 4.0));
    printf("%-28s warps=%2d  cycles per warp-instruction per scheduler: %6.2f\n", name, warps, per_sched);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16}) {
        if (w == 4) printf("-- 1 warp per scheduler (latency-exposed with 8 independent chains)\n");
        if (w == 8) printf("-- 2 warps per scheduler\n");
        if (w == 16) printf("-- 4 warps per scheduler\n");
        run<0>("tanh.approx.f32", w);
        run<1>("ex2.approx.f32", w);
        run<11>("rcp.approx.f32", w);
        run<2>("tanh.approx.f16x2", w);
        run<3>("tanh.approx.bf16x2", w);
        run<4>("fma.rn.f32x2", w);
        run<5>("fma.rn.f32", w);
        run<6>("shfl.bfly.b32", w);
        run<7>("cvt.rn.bf16x2.f32", w);
        run<12>("cvt.rn.f16x2.f32", w);
        run<8>("ld.shared.v4 broadcast", w);
        run<9>("ld.shared.v4 4 addresses", w);
        run<10>("stmatrix.x4", w);
    }
    return 0;
}
