// Probe: register layout of tcgen05.ld.16x256b.x8 (and stmatrix.x4) against a pattern written with tcgen05.st.32x32b.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_layout tmem_layout.cu ; run on a B200
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(uint32_t* out, uint32_t* out_sm) {
    __shared__ uint32_t slot;
    __shared__ __align__(1024) uint8_t stg[4096];
    const int lane = threadIdx.x;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t0 = slot;
    // row = lane, col c: value = row * 256 + c
    for (int c = 0; c < 64; c += 4) {
        uint32_t a = lane * 256 + c, b = a + 1, cc = a + 2, d = a + 3;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t0 + c), "r"(a), "r"(b), "r"(cc), "r"(d) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
            "tcgen05.wait::ld.sync.aligned;"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(t0 + ((uint32_t)(half * 16) << 16))
            : "memory");
        for (int i = 0; i < 32; ++i) out[(half * 32 + lane) * 32 + i] = v[i];
    }
    // stmatrix.x4 probe: register m of lane l = (m << 16 | l) packed as two b16 (lo = 2*l, hi = 2*l+1 within matrix m)
    {
        uint32_t r[4];
        for (int m = 0; m < 4; ++m) r[m] = ((uint32_t)(m * 64 + 2 * lane + 1) << 16) | (uint32_t)(m * 64 + 2 * lane);
        for (int i = lane; i < 1024; i += 32) reinterpret_cast<uint32_t*>(stg)[i] = 0xffffffffu;
        __syncwarp();
        const uint32_t addr = (uint32_t)__cvta_generic_to_shared(stg) + lane * 128;      // lane l -> row l (16 B at column 0)
        asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
        __syncwarp();
        for (int i = lane; i < 1024; i += 32) out_sm[i] = reinterpret_cast<uint32_t*>(stg)[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t0), "r"(64) : "memory");
}

int main() {
    uint32_t *d, *d2;
    cudaMalloc(&d, 64 * 32 * 4);
    cudaMalloc(&d2, 4096);
    probe<<<1, 32>>>(d, d2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    static uint32_t h[64 * 32], h2[1024];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(h2, d2, sizeof(h2), cudaMemcpyDeviceToHost);
    for (int half = 0; half < 2; ++half)
        for (int lane = 0; lane < 32; lane += (lane < 8 ? 1 : 8)) {
            printf("half %d lane %2d:", half, lane);
            for (int i = 0; i < 12; ++i) printf(" (r%d,c%d)", h[(half * 32 + lane) * 32 + i] / 256, h[(half * 32 + lane) * 32 + i] % 256);
            printf(" ... r[28..31]:");
            for (int i = 28; i < 32; ++i) printf(" (r%d,c%d)", h[(half * 32 + lane) * 32 + i] / 256, h[(half * 32 + lane) * 32 + i] % 256);
            printf("\n");
        }
    // check hypothesis: reg 4k+j of lane t in half h = row 16h + t/4 + 8*(j/2), col 8k + 2*(t%4) + (j%2)
    int bad = 0;
    for (int half = 0; half < 2; ++half)
        for (int t = 0; t < 32; ++t)
            for (int i = 0; i < 32; ++i) {
                const int k = i / 4, j = i % 4;
                const uint32_t want = (16 * half + t / 4 + 8 * (j / 2)) * 256 + 8 * k + 2 * (t % 4) + (j % 2);
                if (h[(half * 32 + t) * 32 + i] != want) ++bad;
            }
    printf("16x256b fragment hypothesis mismatches: %d\n", bad);
    // stmatrix: row l of the staging (128 B pitch) should hold matrix l/8 row l%8 = 8 b16 values: element e of that row comes
    // from lane (l%8)*4 + e/2 of register l/8
    bad = 0;
    for (int l = 0; l < 32; ++l)
        for (int e2 = 0; e2 < 4; ++e2) {
            const int m = l / 8, src = (l % 8) * 4 + e2;
            const uint32_t want = ((uint32_t)(m * 64 + 2 * src + 1) << 16) | (uint32_t)(m * 64 + 2 * src);
            if (h2[l * 32 + e2] != want) ++bad;
        }
    printf("stmatrix.x4 hypothesis mismatches: %d\n", bad);
    return 0;
}
