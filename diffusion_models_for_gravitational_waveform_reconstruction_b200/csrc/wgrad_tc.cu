// tcgen05 / TMA weight-gradient GEMM of Conv1d(k=3, pad=1) on bf16 channels-last tensors (sm_100a).
//
//   G_s[m, n] = sum over samples b and rows r of  A[b, r, m] * X[b, r + s, n],   s in {-1, 0, +1}
//
// A = gradient wrt the conv output (d_raw) and X = the conv input, both [B, rows, C] with the channel contiguous.
// The reduction runs over rows, so both MMA operands are "MN-major": a TMA box [64 channels x R rows] with the
// 128-byte swizzle is exactly the canonical MN-major SW128 UMMA layout (8-row groups 1024 B apart = SBO, 64-channel
// blocks one box apart = LBO), no transposed copy of any activation is ever made.  The three taps read ONE X box of
// 64+2 rows through row-shifted descriptors (start address + s*128 B); out-of-range rows (conv zero padding, ragged
// last chunk) are TMA zero fill, and the 3-D tensor maps keep a halo from crossing into the next sample.
//
// Work split: CTA = (128 x BN output tile, split-K share of the (sample, 64-row chunk) list); the three fp32
// accumulators (3 x BN <= 384 TMEM columns) stay resident for the CTA's whole K range and are written once to a
// partial buffer [split][tile][3][128][BN]; gw_wgrad_combine then sums the splits in fixed order into dW[co][ci][k].
// Pipeline: warp 0 TMA producer, warp 1 single-thread MMA issuer, warps 2-5 epilogue (tcgen05.ld -> global).
#include "common.cuh"
#include "../../include/gwb200.h"
#include "tc_common.cuh"
#include <string.h>

#define WG_KR 64                       // rows (K) per pipeline stage
#define WG_A_BOX (WG_KR * 128)         // 8 KB: [64 rows][64 ch] bf16
#define WG_X_BOX 9216                  // [66 rows][64 ch] = 8448 B, padded to 1 KB so every box start is swizzle-aligned
#define WG_X_BYTES ((WG_KR + 2) * 128)

struct WgParams {
    int rows;        // rows per sample (of both views)
    int n_chunk;     // ceil(rows / 64)
    int B;
    int mt, nt;      // output tiles along M (128 each) and N (bn each)
    int bn;          // 64 or 128
    int n_split;
    int stages;
    int lo_tiles;    // mode 1 with Cout >= 128: M tiles [0, lo_tiles) hold only "lo" rows (position 2r) and need taps {-1, 0},
                     // the rest only "hi" rows and need taps {0, +1}; 0 = every tile needs all three taps
};

// MN-major, 128-byte-swizzled operand: LBO [16,30) = byte distance between 64-element MN blocks, SBO [32,46) = between
// 8-row K groups; version 1 [46,48); layout SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t make_mn_sw128_desc(uint32_t saddr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}
// instruction descriptor with both operands MN-major (a_major bit 15, b_major bit 16)
__device__ __forceinline__ uint32_t make_idesc_mn(uint32_t m, uint32_t n) {
    return make_idesc(m, n) | (1u << 15) | (1u << 16);
}

__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_x,
                const __grid_constant__ WgParams P, float* __restrict__ partial, int shifted_desc, float* __restrict__ dW_atomic,
                int mode, int Cout, int Cin_total, int ci_off) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int stages = P.stages;
    const int nb = P.bn / 64;                                      // X boxes per stage
    const int xcopies = shifted_desc ? 1 : 3;                      // fallback: one X box set per tap
    const uint32_t a_bytes = 2 * WG_A_BOX;                         // M = 128 -> two 64-channel boxes
    const uint32_t x_bytes = (uint32_t)xcopies * nb * WG_X_BOX;
    const uint32_t sA = base;
    const uint32_t sX = sA + stages * a_bytes;
    const uint32_t sMisc = sX + stages * x_bytes;
    uint8_t* misc = smem_raw + (sMisc - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(misc);            // full[stages], empty[stages], acc_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 8 * 20);
    auto full_bar = [&](int s) { return smem_u32(bars + s); };
    auto empty_bar = [&](int s) { return smem_u32(bars + stages + s); };
    const uint32_t acc_full = smem_u32(bars + 2 * stages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = P.mt * P.nt;
    const int tile = blockIdx.x % n_tiles, split = blockIdx.x / n_tiles;
    const int m0 = (tile / P.nt) * 128, n0 = (tile % P.nt) * P.bn;
    const int tap_lo = (P.lo_tiles > 0 && tile / P.nt >= P.lo_tiles) ? 1 : 0;
    const int tap_hi = (P.lo_tiles > 0 && tile / P.nt < P.lo_tiles) ? 1 : 2;
    const long total = (long)P.B * P.n_chunk;
    const uint32_t tmem_cols = 512;                                // 3 x bn <= 384, power of two

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_x) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (long w = split; w < total; w += P.n_split) {
                const int b = (int)(w / P.n_chunk), r0 = (int)(w % P.n_chunk) * WG_KR;
                mbar_wait(empty_bar(stage), phase ^ 1);
                mbar_expect_tx(full_bar(stage), a_bytes + (uint32_t)xcopies * nb * (shifted_desc ? WG_X_BYTES : WG_A_BOX));
                tma_load_3d(sA + stage * a_bytes, &tm_a, full_bar(stage), m0, r0, b);
                tma_load_3d(sA + stage * a_bytes + WG_A_BOX, &tm_a, full_bar(stage), m0 + 64, r0, b);
                for (int c = 0; c < xcopies; ++c)
                    for (int j = 0; j < nb; ++j)
                        tma_load_3d(sX + stage * x_bytes + (uint32_t)(c * nb + j) * WG_X_BOX, &tm_x, full_bar(stage), n0 + j * 64,
                                    r0 - 1 + c, b);
                if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint32_t idesc = make_idesc_mn(128, (uint32_t)P.bn);
            const uint64_t a_hi = make_mn_sw128_desc(0, WG_A_BOX), x_hi = make_mn_sw128_desc(0, WG_X_BOX);
            bool first = true;
            for (long w = split; w < total; w += P.n_split) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t a0 = sA + stage * a_bytes, x0 = sX + stage * x_bytes;
#pragma unroll
                for (int tap = 0; tap < 3; ++tap) {
                    if (tap < tap_lo || tap > tap_hi) continue;     // this tile's rows never use that tap (see WgParams::lo_tiles)
                    const uint32_t xt = shifted_desc ? x0 + (uint32_t)tap * 128u : x0 + (uint32_t)tap * nb * WG_X_BOX;
                    // descriptors = constant high word | (address >> 4); a K step of 16 rows is +2048 B = +128 in that field
                    const uint64_t ad0 = a_hi | (uint64_t)(a0 >> 4), xd0 = x_hi | (uint64_t)(xt >> 4);
                    const uint32_t d = tmem_base + (uint32_t)(tap * P.bn);
                    umma_bf16(d, ad0, xd0, idesc, first ? 0u : 1u);
                    umma_bf16(d, ad0 + 128, xd0 + 128, idesc, 1u);
                    umma_bf16(d, ad0 + 256, xd0 + 256, idesc, 1u);
                    umma_bf16(d, ad0 + 384, xd0 + 384, idesc, 1u);
                }
                first = false;
                umma_commit(empty_bar(stage));
                if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(acc_full);
        }
    } else {
        // epilogue: warp q owns TMEM lanes [32q, 32q+32) = output rows m0 + 32q + lane
        const int q = warp & 3;
        const bool any = split < total;                            // a split with no work writes zeros
        if (any) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
        }
        if (dW_atomic != nullptr) {
            // split-K by fp32 atomics straight into dW (red.global.add.f32): no partial buffer, no fold / scatter passes;
            // the summation order over the CTAs is not fixed (run-to-run differences at the 1e-7 level)
            const int m = m0 + q * 32 + lane;                        // row of the [M, N] product this lane owns
            const int M = mode == 1 ? 2 * Cout : Cout;
            for (int tap = tap_lo; tap <= tap_hi && any; ++tap) {
                for (int c0 = 0; c0 < P.bn; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tap * P.bn + c0), v);
                    if (m >= M) continue;
                    const int co = mode == 1 ? m % Cout : m;
                    const bool hi = mode == 1 && m >= Cout;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float g = __uint_as_float(v[i]);
                        float* d = dW_atomic + ((size_t)co * Cin_total + ci_off + n0 + c0 + i) * 3;
                        if (mode == 0) {
                            atomicAdd(d + tap, g);
                        } else if (!hi) {                            // lo rows (position 2r): G_-1 -> k0, G_0 -> k1 and k2
                            if (tap == 0) atomicAdd(d + 0, g);
                            else if (tap == 1) { atomicAdd(d + 1, g); atomicAdd(d + 2, g); }
                        } else {                                     // hi rows (position 2r+1): G_0 -> k0 and k1, G_+1 -> k2
                            if (tap == 1) { atomicAdd(d + 0, g); atomicAdd(d + 1, g); }
                            else if (tap == 2) atomicAdd(d + 2, g);
                        }
                    }
                }
            }
        } else {
        float* out = partial + ((size_t)(split * n_tiles + tile) * 3) * 128 * P.bn;
        for (int tap = 0; tap < 3; ++tap) {
            for (int c0 = 0; c0 < P.bn; c0 += 32) {
                uint32_t v[32];
                if (any && tap >= tap_lo && tap <= tap_hi) {
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tap * P.bn + c0), v);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0u;
                }
                float4* dst = reinterpret_cast<float4*>(out + ((size_t)tap * 128 + q * 32 + lane) * P.bn + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                         __uint_as_float(v[4 * i + 3]));
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------ combine
// stage 1: fold the split-K partials [n_split][cols] into row 0 in fixed order (32 columns x 32 split lanes per block);
// stage 2: scatter the folded G_s[m][n] into dW[co][ci][k]:
//   mode 0 (plain):   dW[m][ci_off + n][k] += G_{k-1}[m][n]                                        (m < Cout)
//   mode 1 (up/pair): A is the pair view [rows = L/2][2*Cout] of d_raw (lo = position 2r, hi = 2r+1) and X = h:
//     dW[co][ci_off + n][0] += G_-1[lo] + G_0[hi];  [1] += G_0[lo] + G_0[hi];  [2] += G_0[lo] + G_+1[hi]
__global__ void __launch_bounds__(256) wgrad_fold_kernel(float* __restrict__ partial, int n_split, long cols) {
    pdl_wait();
    pdl_launch_dependents();
    // block = 32 float4 columns x 8 split lanes; each lane streams its splits with 4 loads in flight
    __shared__ float4 red[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long c4 = (long)blockIdx.x * 32 + tx;            // float4 column
    const long n4 = cols / 4;
    float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (c4 < n4) {
        const float4* p = reinterpret_cast<const float4*>(partial) + c4;
#pragma unroll 4
        for (int r = ty; r < n_split; r += 8) {
            const float4 v = p[(size_t)r * n4];
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
    }
    red[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && c4 < n4) {
        float4 sacc = red[0][tx];
#pragma unroll
        for (int t = 1; t < 8; ++t) {
            const float4 v = red[t][tx];
            sacc.x += v.x; sacc.y += v.y; sacc.z += v.z; sacc.w += v.w;
        }
        reinterpret_cast<float4*>(partial)[c4] = sacc;
    }
}

__global__ void __launch_bounds__(256) wgrad_scatter_kernel(const float* __restrict__ folded, WgParams P, int mode, int Cout,
                                                            int Cx, int Cin_total, int ci_off, float* __restrict__ dW) {
    pdl_wait();
    pdl_launch_dependents();
    const long n = (long)Cout * Cx;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int co = (int)(i / Cx), nn = (int)(i % Cx);
    const int tn = nn / P.bn, cn = nn % P.bn;
    auto G = [&](int m, int tap) {
        const int tile = (m / 128) * P.nt + tn;
        return folded[(((size_t)tile * 3 + tap) * 128 + (m % 128)) * P.bn + cn];
    };
    float* d = dW + ((size_t)co * Cin_total + ci_off + nn) * 3;
    if (mode == 0) {
        d[0] += G(co, 0);
        d[1] += G(co, 1);
        d[2] += G(co, 2);
    } else {
        const float lo0 = G(co, 1), hi0 = G(Cout + co, 1);
        d[0] += G(co, 0) + hi0;
        d[1] += lo0 + hi0;
        d[2] += lo0 + G(Cout + co, 2);
    }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiledW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledW wg_get_encode() {
    static PFN_encodeTiledW fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiledW)p;
    }
    return fn;
}
static int wg_map3(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1) {
    PFN_encodeTiledW enc = wg_get_encode();
    GW_REQUIRE(enc != nullptr, "wgrad_tc: cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GW_REQUIRE(r == CUDA_SUCCESS, "wgrad_tc: cuTensorMapEncodeTiled failed with %d (dims %llu %llu %llu)", (int)r,
               (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2);
    return GW_OK;
}

static int wg_params(int mode, int B, int L, int Cout, int Cx, WgParams* P) {
    GW_REQUIRE(mode == 0 || mode == 1, "wgrad_tc: mode %d", mode);
    GW_REQUIRE(Cout % 64 == 0 && Cx % 64 == 0 && Cout > 0 && Cx > 0, "wgrad_tc: Cout=%d Cx=%d must be multiples of 64", Cout, Cx);
    GW_REQUIRE(mode == 0 || L % 2 == 0, "wgrad_tc: the upsampled source needs an even length");
    memset(P, 0, sizeof(*P));
    const int M = mode == 1 ? 2 * Cout : Cout;
    P->rows = mode == 1 ? L / 2 : L;
    P->n_chunk = gw_cdiv(P->rows, WG_KR);
    P->B = B;
    P->mt = gw_cdiv(M, 128);                    // M = 64: the upper half of the tile is TMA zero fill
    P->bn = Cx % 128 == 0 ? 128 : 64;
    P->nt = Cx / P->bn;
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles = P->mt * P->nt;
    long ns = sms / tiles;
    if (ns < 1) ns = 1;
    const long total = (long)B * P->n_chunk;
    if (ns > total) ns = total;
    P->n_split = (int)ns;
    P->lo_tiles = (mode == 1 && Cout % 128 == 0) ? Cout / 128 : 0;
    return GW_OK;
}

extern "C" long gw_wgrad_tc_scratch_elems(int mode, int B, int L, int Cout, int Cx) {
    WgParams P;
    if (wg_params(mode, B, L, Cout, Cx, &P) != GW_OK) return -1;
    return (long)P.n_split * P.mt * P.nt * 3 * 128 * P.bn;
}

// d_raw [B, L, Cout] bf16; x [B, Lx, Cx] bf16 with Lx = L (mode 0) or L/2 (mode 1: x is h before the nearest upsample);
// dW fp32 [Cout][Cin_total][3], this call ACCUMULATES the block of input channels [ci_off, ci_off + Cx).
// variant bit 0: do not use row-shifted descriptors (load one X box per tap instead);
// variant bit 1: split-K by fp32 atomics into dW instead of the deterministic partial-buffer + fold + scatter passes.
extern "C" int gw_wgrad_tc(int mode, const void* d_raw, const void* x, int B, int L, int Cout, int Cx, int Cin_total, int ci_off,
                           float* scratch, long scratch_elems, float* dW, int variant, void* stream) {
    WgParams P;
    int rc = wg_params(mode, B, L, Cout, Cx, &P);
    if (rc != GW_OK) return rc;
    GW_REQUIRE(d_raw != nullptr && x != nullptr && scratch != nullptr && dW != nullptr, "wgrad_tc: null pointer");
    GW_REQUIRE(ci_off >= 0 && ci_off + Cx <= Cin_total, "wgrad_tc: channel block out of range");
    const long need = (long)P.n_split * P.mt * P.nt * 3 * 128 * P.bn;
    GW_REQUIRE(need <= scratch_elems, "wgrad_tc: scratch too small (%ld < %ld)", scratch_elems, need);
    const int shifted = (variant & 1) ? 0 : 1;
    CUtensorMap ta, tx;
    const uint64_t Ma = mode == 1 ? 2 * (uint64_t)Cout : (uint64_t)Cout;
    if ((rc = wg_map3(&ta, d_raw, Ma, (uint64_t)P.rows, (uint64_t)B, 64, WG_KR)) != GW_OK) return rc;
    if ((rc = wg_map3(&tx, x, (uint64_t)Cx, (uint64_t)P.rows, (uint64_t)B, 64, shifted ? WG_KR + 2 : WG_KR)) != GW_OK) return rc;
    const int nb = P.bn / 64;
    const int stage_bytes = 2 * WG_A_BOX + (shifted ? 1 : 3) * nb * WG_X_BOX;
    int stages = (232448 - 1024 - 512) / stage_bytes;
    if (stages > 8) stages = 8;
    P.stages = stages;
    const int smem = 1024 + stages * stage_bytes + 512;
    GW_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaStream_t st = (cudaStream_t)stream;
    const bool atomic = (variant & 2) != 0;
    GW_CUDA(gw_launch_pdl(wgrad_tc_kernel, P.mt * P.nt * P.n_split, dim3(192), (size_t)(smem), st, ta, tx, P, scratch, shifted, atomic ? dW : nullptr, mode, Cout,
                                                                 Cin_total, ci_off));
    GW_LAUNCH_CHECK();
    if (atomic || (variant & 4)) return GW_OK;               // bit 2: the caller runs gw_wgrad_tc_finish (e.g. on another stream)
    const long cols = (long)P.mt * P.nt * 3 * 128 * P.bn;
    GW_CUDA(gw_launch_pdl(wgrad_fold_kernel, dim3((unsigned)((cols / 4 + 31) / 32)), dim3(256), (size_t)(0), st, scratch, P.n_split, cols));
    GW_LAUNCH_CHECK();
    const long n = (long)Cout * Cx;
    GW_CUDA(gw_launch_pdl(wgrad_scatter_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), st, scratch, P, mode, Cout, Cx, Cin_total, ci_off, dW));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// the deterministic split-K epilogue of gw_wgrad_tc(variant | 4): fold the partial buffer, then scatter-accumulate into dW.  Same
// shape arguments as the GEMM call that filled `scratch`; may run on any stream ordered after that call.
extern "C" int gw_wgrad_tc_finish(int mode, int B, int L, int Cout, int Cx, int Cin_total, int ci_off, float* scratch, float* dW,
                                  void* stream) {
    WgParams P;
    int rc = wg_params(mode, B, L, Cout, Cx, &P);
    if (rc != GW_OK) return rc;
    GW_REQUIRE(scratch != nullptr && dW != nullptr, "wgrad_tc_finish: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const long cols = (long)P.mt * P.nt * 3 * 128 * P.bn;
    GW_CUDA(gw_launch_pdl(wgrad_fold_kernel, dim3((unsigned)((cols / 4 + 31) / 32)), dim3(256), (size_t)(0), st, scratch, P.n_split, cols));
    GW_LAUNCH_CHECK();
    const long n = (long)Cout * Cx;
    GW_CUDA(gw_launch_pdl(wgrad_scatter_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), st, scratch, P, mode, Cout, Cx, Cin_total, ci_off, dW));
    GW_LAUNCH_CHECK();
    return GW_OK;
}
