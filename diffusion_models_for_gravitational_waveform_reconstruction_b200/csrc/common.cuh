// Shared device/host helpers for the gwb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#define GW_OK 0
#define GW_ERR_ARG -1
#define GW_ERR_CUDA -2
#define GW_ERR_UNSUPPORTED -3

// ---------------------------------------------------------------- error plumbing
void gw_set_error(const char* fmt, ...);

#define GW_REQUIRE(cond, ...)                                  \
    do {                                                       \
        if (!(cond)) {                                         \
            gw_set_error(__VA_ARGS__);                         \
            return GW_ERR_ARG;                                 \
        }                                                      \
    } while (0)

#define GW_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            gw_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return GW_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define GW_LAUNCH_CHECK()                                                               \
    do {                                                                                \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess) {                                                        \
            gw_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return GW_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

typedef __nv_bfloat16 bf16;

// dtype codes of the C-ABI
#define GW_F32 0
#define GW_BF16 1
#define GW_DOTS 2     // gw_final_step only: h = head dot products [Bn, L, 4] fp32 (gw_conv_gn2)

static inline int gw_cdiv(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- 8-wide vector load/store
// Activations are channels-last; a thread always owns 8 consecutive channels.
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
    uint4 r;
    r.x = pack_bf16x2(v[0], v[1]);
    r.y = pack_bf16x2(v[2], v[3]);
    r.z = pack_bf16x2(v[4], v[5]);
    r.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = r;
}
// value as it will be read back from storage (stats must describe the stored data)
__device__ __forceinline__ float round_to(float x, const float*) { return x; }
__device__ __forceinline__ float round_to(float x, const bf16*) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }

// ---------------------------------------------------------------- packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2, sm_100a)
// Issue-bound elementwise / epilogue code halves its FP instruction count with these; a value is two floats in a b64.
typedef unsigned long long f32x2;
__device__ __forceinline__ unsigned long long pk2(uint32_t lo, uint32_t hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float& lo, float& hi) {
    uint32_t a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
    lo = __uint_as_float(a);
    hi = __uint_as_float(b);
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long pkf2(float lo, float hi) { return pk2(__float_as_uint(lo), __float_as_uint(hi)); }

// ---------------------------------------------------------------- activations
template <bool FAST>
__device__ __forceinline__ float sigmoid_f(float x) {
    if (FAST) return __fdividef(1.0f, 1.0f + __expf(-x));
    return 1.0f / (1.0f + expf(-x));
}
template <bool FAST>
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f<FAST>(x); }

// FAST silu: x*sigmoid(x) = h + h*tanh(h), h = x/2; one MUFU op (tanh.approx, rel. error ~2^-11 << bf16 epsilon)
__device__ __forceinline__ float silu_tanh(float x) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Sum V per-lane values over the 32 lanes with recursive halving (V-1 + 5-log2(V) shuffles instead of 5V): afterwards the
// total of value j sits in every lane whose top log2(V) lane bits equal j.
template <int V>
__device__ __forceinline__ float warp_reduce_multi(float (&v)[V], int lane) {
    static_assert(V == 4 || V == 8 || V == 16, "V");
    int n = V;
#pragma unroll
    for (int m = 16; n > 1; m >>= 1) {
        n >>= 1;
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < V / 2; ++i) {
            if (i < n) {
                const float send = up ? v[i] : v[i + n];
                const float keep = up ? v[i + n] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
            }
        }
    }
    float t = v[0];
    constexpr int REST = V == 16 ? 1 : (V == 8 ? 2 : 4);      // lane bits not consumed by the halving
#pragma unroll
    for (int m = REST; m > 0; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
    return t;
}

// Philox4x32-10 counter RNG + Box-Muller; key = seed, counter = (sample, step, chunk, 0).
struct Philox {
    __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
        uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    __device__ static inline void gen(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) {
        uint32_t c[4] = {c0, c1, c2, c3};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            round(c, k0, k1);
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
    // 4 standard normals for (sample, step, chunk): positions 4*chunk .. 4*chunk+3
    __device__ static inline void normal4(uint64_t seed, uint32_t sample, uint32_t step, uint32_t chunk, float (&z)[4]) {
        uint32_t r[4];
        gen(seed, chunk, step, sample, 0x6a09e667u, r);
        const float k = 2.3283064365386963e-10f;  // 2^-32
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float u1 = ((float)r[2 * i] + 0.5f) * k;   // (0,1)
            float u2 = ((float)r[2 * i + 1] + 0.5f) * k;
            u1 = fminf(fmaxf(u1, 1.0e-10f), 1.0f);
            float rad = sqrtf(-2.0f * logf(u1));
            float s, c;
            sincosf(6.283185307179586f * u2, &s, &c);
            z[2 * i] = rad * c;
            z[2 * i + 1] = rad * s;
        }
    }
};

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): consecutive kernels of the sampler's reverse step are launched with the programmatic
// stream-serialization attribute, so a kernel's CTAs may start (as SMs free up) and run their prologue -- barrier init, TMEM
// allocation, weight staging -- while the previous kernel drains; pdl_wait() blocks until the previous grid has completed and its
// writes are visible, and must precede the first access to anything the previous kernel wrote.  Both are no-ops in a kernel
// launched without the attribute.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern int g_pdl;      // forward.cu; gw_set_option("pdl", 0/1)
template <typename... KArgs, typename... Args>
static inline cudaError_t gw_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// the same with thread-block clusters of `cluster` consecutive CTAs (1 = none); cluster = 2 is a tcgen05 CTA pair
template <typename... KArgs, typename... Args>
static inline cudaError_t gw_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                                            bool force_pdl, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (g_pdl || force_pdl) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = (unsigned)cluster;
        at[n].val.clusterDim.y = 1;
        at[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
