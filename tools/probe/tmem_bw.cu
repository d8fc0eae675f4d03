// Probe: tcgen05.ld throughput (bytes per cycle per SM) for the shapes the epilogues use, with 4 and 8 warps, and whether a
// warp streaming TMEM loads slows an FMA-bound warp on the same scheduler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LD32(addr)                                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"              \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),   \
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),  \
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])               \
                 : "r"(addr) : "memory")
#define LD16x8(addr)                                                                                                    \
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"              \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),   \
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),  \
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])               \
                 : "r"(addr) : "memory")

// mode 0: 32x32b.x32 (32 lanes x 32 cols = 4 KB / message); 1: 16x256b.x8 (16 lanes x 64 cols = 4 KB / message)
// mix: warps >= 4 run an FMA loop instead of loads
template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, int iters, int mix) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t0a = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t v[32];
    uint32_t acc = 0;
    float f[8];
    for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 0.01f + i;
    __syncthreads();
    const long long t0 = clock64();
    if (mix && warp >= 4) {
        for (int it = 0; it < iters * 16; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f[i]));
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (MODE == 0) LD32(t0a + (uint32_t)(c * 32 + (warp >> 2) * 256) % 512);
                else LD16x8(t0a + (uint32_t)((c & 3) * 64 + (warp >> 2) * 256) % 512 + ((uint32_t)((c >> 2) * 16) << 16));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += v[0] + v[31];
            }
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = acc + (uint32_t)f[0];
    cyc[warp] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int MODE>
void run(const char* name, int warps, int mix) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8 * 32);
    const int iters = 256;
    k<MODE><<<1, warps * 32>>>(out, cyc, iters, mix);
    k<MODE><<<1, warps * 32>>>(out, cyc, iters, mix);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[32]; cudaMemcpy(c, cyc, 8 * warps, cudaMemcpyDeviceToHost);
    const int ld_warps = mix ? 4 : warps;
    const double bytes = (double)ld_warps * iters * 8 * 4096;
    printf("%-16s warps=%d mix=%d (%s): load warps %lld cycles -> %.1f B/clk/SM", name, warps, mix, cudaGetErrorString(e), c[0], bytes / (double)c[0]);
    if (mix) printf("; FMA warps %lld cycles for %d FMAs each -> %.2f cyc/FMA/scheduler", c[4], iters * 16 * 8, (double)c[4] / (iters * 16 * 8));
    printf("\n");
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("32x32b.x32", 4, 0);
    run<0>("32x32b.x32", 8, 0);
    run<1>("16x256b.x8", 4, 0);
    run<1>("16x256b.x8", 8, 0);
    run<0>("32x32b.x32", 8, 1);
    run<1>("16x256b.x8", 8, 1);
    return 0;
}
