// Forward-path kernels other than the tcgen05 conv: FiLM vectors, conditioning pyramid, first conv,
// exact-mode SIMT conv, fused GroupNorm-apply/SiLU/cond/FiLM/pool, fused head conv + DDIM/DDPM step, q_sample.
// Reference call sites are cited in include/gwb200.h.
#include "common.cuh"
#include "../../include/gwb200.h"
#include <string.h>

// ------------------------------------------------------------------------------------------------
// error text
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "ok";
void gw_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* gw_last_error(void) { return g_err; }
extern "C" int gw_version(void) { return 100; }
extern "C" int gw_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    GW_CUDA(cudaGetDevice(&dev));
    GW_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    GW_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    GW_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------
// FiLM vectors: one CTA per FILM_NB timestep values, so every weight row is read once per FILM_NB samples
// ------------------------------------------------------------------------------------------------
#define FILM_NB 4
__global__ void __launch_bounds__(256) film_kernel(const int64_t* __restrict__ t, int n, int time_dim, float inv_max_time_den,
                                                   float freq_coef, const float* __restrict__ w1,
                                                   const float* __restrict__ b1, const float* __restrict__ w2,
                                                   const float* __restrict__ b2, int base, int F, float* __restrict__ out,
                                                   float* __restrict__ aux) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(16) float sm[];
    float* emb = sm;                         // [NB][time_dim]
    float* act = sm + FILM_NB * time_dim;    // [NB][base]
    const int n0 = blockIdx.x * FILM_NB;
    const int half = time_dim / 2;
    const int na = time_dim + 3 * base;
    for (int idx = threadIdx.x; idx < FILM_NB * time_dim; idx += blockDim.x) {
        const int s = idx / time_dim, i = idx % time_dim;
        float v = 0.0f;
        if (n0 + s < n && i < 2 * half) {
            const float ts = (float)t[n0 + s] / inv_max_time_den;
            const int j = i < half ? i : i - half;
            const float a = ts * expf((float)j * freq_coef);
            v = i < half ? sinf(a) : cosf(a);
        }
        emb[idx] = v;
        if (aux != nullptr && blockIdx.y == 0 && n0 + s < n) aux[(size_t)(n0 + s) * na + i] = v;
    }
    __syncthreads();
    // time_mlp Linear: one warp per output row (coalesced weight reads), FILM_NB dot products at once
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = warp; j < base; j += 8) {
        float acc[FILM_NB];
#pragma unroll
        for (int s = 0; s < FILM_NB; ++s) acc[s] = 0.0f;
        const float* wr = w1 + (size_t)j * time_dim;
        for (int i = lane; i < time_dim; i += 32) {
            const float w = wr[i];
#pragma unroll
            for (int s = 0; s < FILM_NB; ++s) acc[s] = fmaf(w, emb[s * time_dim + i], acc[s]);
        }
#pragma unroll
        for (int s = 0; s < FILM_NB; ++s) acc[s] = warp_sum(acc[s]);
        if (lane < FILM_NB) {
            float pre = acc[0];
#pragma unroll
            for (int s = 1; s < FILM_NB; ++s) pre = lane == s ? acc[s] : pre;
            pre += b1[j];
            const float ctx = silu_f<false>(pre);     // time_mlp's SiLU
            const float a2 = silu_f<false>(ctx);      // tproj's leading SiLU
            act[lane * base + j] = a2;
            if (aux != nullptr && blockIdx.y == 0 && n0 + lane < n) {    // saved for gw_film_bwd: [emb | pre | ctx | act]
                float* ax = aux + (size_t)(n0 + lane) * na + time_dim;
                ax[j] = pre;
                ax[base + j] = ctx;
                ax[2 * base + j] = a2;
            }
        }
    }
    __syncthreads();
    // tproj Linears: one thread per output, its weight row streamed as float4; blockIdx.y splits the F outputs so that the grid
    // covers the GPU at small n (the time MLP above is recomputed per split: 8 K MACs per sample)
    const int f_per = (F + gridDim.y - 1) / gridDim.y, f_end = min(F, (int)(blockIdx.y + 1) * f_per);
    for (int f = blockIdx.y * f_per + threadIdx.x; f < f_end; f += blockDim.x) {
        float acc[FILM_NB];
#pragma unroll
        for (int s = 0; s < FILM_NB; ++s) acc[s] = 0.0f;
        const float4* wr = reinterpret_cast<const float4*>(w2 + (size_t)f * base);
        for (int j4 = 0; j4 < base / 4; ++j4) {
            const float4 w = wr[j4];
#pragma unroll
            for (int s = 0; s < FILM_NB; ++s) {
                const float4 a = *reinterpret_cast<const float4*>(act + s * base + j4 * 4);
                acc[s] = fmaf(w.x, a.x, fmaf(w.y, a.y, fmaf(w.z, a.z, fmaf(w.w, a.w, acc[s]))));
            }
        }
        const float bb = b2[f];
#pragma unroll
        for (int s = 0; s < FILM_NB; ++s)
            if (n0 + s < n) out[(size_t)(n0 + s) * F + f] = acc[s] + bb;
    }
}

extern "C" int gw_film_vectors(const int64_t* t, int n, int time_dim, float max_time, const float* w1, const float* b1,
                               const float* w2, const float* b2, int base, int F, float* out, float* aux, void* stream) {
    GW_REQUIRE(n > 0 && time_dim > 0 && base > 0 && F > 0 && base % 4 == 0, "gw_film_vectors: bad sizes (base must be a multiple of 4)");
    const int half = time_dim / 2;
    const float den = max_time > 1.0f ? max_time : 1.0f;                       // models.py:21
    const float coef = (float)(-(log(10000.0) / (double)(half - 1 > 1 ? half - 1 : 1)));   // models.py:25
    size_t smem = (size_t)FILM_NB * (time_dim + base) * sizeof(float);
    const int nbx = gw_cdiv(n, FILM_NB);
    int f_split = nbx >= 296 ? 1 : gw_cdiv(296, nbx);                          // ~2 CTAs per SM
    if (f_split > gw_cdiv(F, 256)) f_split = gw_cdiv(F, 256);
    GW_CUDA(gw_launch_pdl(film_kernel, dim3(nbx, f_split), dim3(256), (size_t)(smem), (cudaStream_t)stream, t, n, time_dim, den, coef, w1, b1, w2, b2, base, F, out, aux));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------
// conditioning pyramid
// ------------------------------------------------------------------------------------------------
struct PyrArgs {
    float* out[GW_MAX_LEVELS];
    int len[GW_MAX_LEVELS];
};

__global__ void __launch_bounds__(256) cond_pyramid_kernel(const float* __restrict__ x, int B, int Cx, int L, int Cc,
                                                           PyrArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    const int lvl = blockIdx.y;
    const int Lo = a.len[lvl];
    const long total = (long)B * Lo;
    const float scale = (float)L / (float)Lo;     // area_pixel_compute_scale (align_corners=False, size given)
    // one thread per (b, l): reads are coalesced along l for every channel, the Cc outputs of a row are contiguous
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int l = (int)(i % Lo);
        const int b = (int)(i / Lo);
        float src = scale * ((float)l + 0.5f) - 0.5f;
        if (src < 0.0f) src = 0.0f;
        int i0 = (int)src;
        if (i0 > L - 1) i0 = L - 1;
        const int i1 = i0 + (i0 < L - 1 ? 1 : 0);
        const float l1 = src - (float)i0;
        const float l0 = 1.0f - l1;
        const float* xp = x + ((size_t)b * Cx + 1) * L;
        float* op = a.out[lvl] + (size_t)i * Cc;
        for (int c = 0; c < Cc; ++c) op[c] = l0 * xp[(size_t)c * L + i0] + l1 * xp[(size_t)c * L + i1];
    }
}

// Power-of-two pyramids (every level is L >> j, the reference's 4096 -> 2048 -> 1024 -> 512): the interpolation is the identity at
// level 0 and the mean of the two central samples of each stride-2^j cell below it (src = 2^j (l + 1/2) - 1/2 lands exactly
// between two samples, weights 1/2 and 1/2; same values as the generic kernel).  A CTA stages 1024 positions of all cond channels
// once (vector loads) and writes every level's rows as one contiguous run: the generic kernel's strided reads and 20-byte row
// stores took 52 us for 23 MB at B = 256.
#define PYR_TILE 1024
__global__ void __launch_bounds__(256) cond_pyramid_pow2_kernel(const float* __restrict__ x, int Cx, int L, int Cc, int n_levels,
                                                                PyrArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(16) float xs[];          // [Cc][PYR_TILE]
    const int b = blockIdx.y, l0 = blockIdx.x * PYR_TILE;
    const float* xp = x + ((size_t)b * Cx + 1) * L + l0;
    for (int i = threadIdx.x; i < Cc * (PYR_TILE / 4); i += 256) {
        const int c = i / (PYR_TILE / 4), q = i % (PYR_TILE / 4);
        *reinterpret_cast<float4*>(xs + c * PYR_TILE + 4 * q) = *reinterpret_cast<const float4*>(xp + (size_t)c * L + 4 * q);
    }
    __syncthreads();
    for (int lvl = 0; lvl < n_levels; ++lvl) {
        const int n_out = PYR_TILE >> lvl, s = 1 << lvl;
        float* op = (lvl == 0 ? a.out[0] : lvl == 1 ? a.out[1] : lvl == 2 ? a.out[2] : a.out[3]) +
                    ((size_t)b * (L >> lvl) + (l0 >> lvl)) * Cc;
        for (int idx = threadIdx.x; idx < n_out * Cc; idx += 256) {
            const int i = idx / Cc, c = idx - i * Cc;
            const float* xr = xs + c * PYR_TILE;
            op[idx] = lvl == 0 ? xr[i] : 0.5f * xr[s * i + (s >> 1) - 1] + 0.5f * xr[s * i + (s >> 1)];
        }
    }
}

extern "C" int gw_cond_pyramid(const float* x, int B, int Cx, int L, int Cc, int n_levels, const int* level_len,
                               float* const* level_out, void* stream) {
    GW_REQUIRE(n_levels > 0 && n_levels <= GW_MAX_LEVELS, "gw_cond_pyramid: n_levels %d", n_levels);
    GW_REQUIRE(Cc > 0 && 1 + Cc <= Cx, "gw_cond_pyramid: Cc %d Cx %d", Cc, Cx);
    PyrArgs a;
    for (int i = 0; i < n_levels; ++i) {
        a.out[i] = level_out[i];
        a.len[i] = level_len[i];
    }
    bool pow2 = L % PYR_TILE == 0 && n_levels <= 4;
    for (int i = 0; i < n_levels; ++i) pow2 = pow2 && level_len[i] == (L >> i);
    if (pow2) {
        const size_t smem = (size_t)Cc * PYR_TILE * sizeof(float);
        GW_CUDA(gw_launch_pdl(cond_pyramid_pow2_kernel, dim3(L / PYR_TILE, B), dim3(256), (size_t)(smem), (cudaStream_t)stream, x, Cx, L, Cc, n_levels, a));
        GW_LAUNCH_CHECK();
        return GW_OK;
    }
    long total = (long)B * L;
    int gx = (int)((total + 255) / 256);
    if (gx > 148 * 16) gx = 148 * 16;
    GW_CUDA(gw_launch_pdl(cond_pyramid_kernel, dim3(gx, n_levels), dim3(256), (size_t)(0), (cudaStream_t)stream, x, B, Cx, L, Cc, a));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------
// deterministic per-tile GroupNorm partials: red[256][2] in smem -> part[(tile, g)]
// each thread contributed (s1, s2) for group `g_of_thread`; groups in this tile = n_groups starting at g0.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_stats_reduce(float s1, float s2, int g_local, int n_groups, float* red,
                                                  float* part_tile /* [8][2] */, int g0) {
    // red layout: [thread][3] = (s1, s2, group)
    red[threadIdx.x * 3 + 0] = s1;
    red[threadIdx.x * 3 + 1] = s2;
    red[threadIdx.x * 3 + 2] = __int_as_float(g_local);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = warp; g < n_groups; g += (blockDim.x >> 5)) {
        float a1 = 0.0f, a2 = 0.0f;
        for (int i = lane; i < (int)blockDim.x; i += 32) {
            if (__float_as_int(red[i * 3 + 2]) == g) {
                a1 += red[i * 3 + 0];
                a2 += red[i * 3 + 1];
            }
        }
        a1 = warp_sum(a1);
        a2 = warp_sum(a2);
        if (lane == 0) {
            part_tile[(g0 + g) * 2 + 0] = a1;
            part_tile[(g0 + g) * 2 + 1] = a2;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// first conv (K1): NCL fp32 input -> channels-last raw + stats.  128 positions x C channels per CTA.
// A thread owns one channel octet for 4 CONSECUTIVE positions: the 6 input samples they need are two vector smem
// loads per input channel and every weight vector is reused by 4 positions (96 FMA per 8 shared-memory loads).
// Lane = (octet, position group) so a warp's store of one position offset covers whole 128-byte rows.
// ------------------------------------------------------------------------------------------------
#define CIN_NPB 4                                 // 128-position blocks per CTA (weights / bias staged once for all of them)
#define CIN_MAX_CC 8
// Arguments of the fused first block (MODE 2): GroupNorm-apply + SiLU + cond 1x1 conv + FiLM + pool on the recomputed conv
struct ConvInApply {
    const float* part_in;     // [B, n_part, 8, 2] partial sums written by the MODE 1 pass
    const float* gn_w;
    const float* gn_b;
    const float* wc;          // [C, Cc]
    const float* bc;          // [C]
    const float* film;        // FiLM rows; this block's (gamma | beta) at film_off
    long film_b_stride, film_step_stride;
    int film_off, Cc;
    void* out;                // [B, L, C]
    void* pooled;             // [B, L/2, C] or NULL
};

// MODE 0: raw conv output + GroupNorm partial sums (training keeps raw for the backward pass)
// MODE 1: partial sums only (nothing but 64 B per tile is written)
// MODE 2: recompute the conv and apply the rest of the block -> out (+ pooled); the 134 MB raw tensor never exists.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) conv_in_kernel(const float* __restrict__ xa, const float* __restrict__ xb,
                                                      const int* __restrict__ step_ptr, int Cx, int L,
                                                      const float* __restrict__ w, const float* __restrict__ bias, int C,
                                                      T* __restrict__ raw, float* __restrict__ part, int n_part, ConvInApply ap) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int TP = 128, XP = CIN_NPB * TP + 8;   // XP: row pitch of xs (multiple of 4 -> float4-aligned groups)
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                              // [Cx][XP]: xs[c][j] = x[c][l00 - 1 + j]
    float* ws = xs + Cx * XP;                    // [Cx*3][C]
    float* bs = ws + Cx * 3 * C;                 // [C]
    float* wst = bs + C;                         // [C/8 octets][8 warps][2]
    float* cf = wst + (C / 8) * 8 * 2;           // MODE 2: [4 + Cc][C] coefficient rows (A, Bn, G, E', W'_j), split-octet layout
    __shared__ float s_mean[8], s_rstd[8];
    const int b = blockIdx.y, l00 = blockIdx.x * (CIN_NPB * TP);
    const int step = step_ptr != nullptr ? *step_ptr : 0;
    const float* x = (step & 1) ? xb : xa;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < Cx * XP; i += blockDim.x) {
        const int c = i / XP, p = i % XP;
        const int l = l00 + p - 1;
        xs[i] = (p < CIN_NPB * TP + 2 && l >= 0 && l < L) ? x[((size_t)b * Cx + c) * L + l] : 0.0f;
    }
    // weights as [ck][half][octet][4]: the 8 octet lanes of a quarter-warp read 8 consecutive float4 (no bank conflicts)
    auto split = [C](int co) { return ((co >> 2) & 1) * (C / 2) + (co >> 3) * 4 + (co & 3); };
    for (int i = threadIdx.x; i < Cx * 3 * C; i += blockDim.x) {
        const int co = i % C, ck = i / C;        // ck = ci*3 + k
        ws[ck * C + split(co)] = w[(size_t)co * Cx * 3 + ck];
    }
    for (int i = threadIdx.x; i < C; i += blockDim.x) bs[i] = bias[i];
    if (MODE == 2) {
        // finish the GroupNorm statistics of this sample (biased variance, eps 1e-5; models.py:154-158)
        const int cg = C / 8;
        if (warp < 8) {
            double a1 = 0.0, a2 = 0.0;
            const float* pp = ap.part_in + (size_t)b * n_part * 16 + warp * 2;
            for (int i = lane; i < n_part; i += 32) {
                a1 += (double)pp[(size_t)i * 16];
                a2 += (double)pp[(size_t)i * 16 + 1];
            }
            a1 = warp_sum_d(a1);
            a2 = warp_sum_d(a2);
            if (lane == 0) {
                const double n = (double)cg * (double)L;
                const double mean = a1 / n;
                double var = a2 / n - mean * mean;
                if (var < 0.0) var = 0.0;
                s_mean[warp] = (float)mean;
                s_rstd[warp] = (float)(1.0 / sqrt(var + 1e-5));
            }
        }
        __syncthreads();
        const float* fr = ap.film + (size_t)step * ap.film_step_stride + (size_t)b * ap.film_b_stride + ap.film_off;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const int g = c / cg, sc = split(c);
            const float a = s_rstd[g] * ap.gn_w[c];
            const float G = 1.0f + fr[c];
            cf[0 * C + sc] = a;
            cf[1 * C + sc] = ap.gn_b[c] - s_mean[g] * a;
            cf[2 * C + sc] = G;
            cf[3 * C + sc] = ap.Cc > 0 ? fmaf(ap.bc[c], G, fr[C + c]) : fr[C + c];       // (h + bc) G + beta, folded
            for (int j = 0; j < ap.Cc; ++j) cf[(4 + j) * C + sc] = ap.wc[c * ap.Cc + j] * G;
        }
    }
    __syncthreads();
    const int pg = warp * 4 + (lane >> 3);       // position group: positions 4*pg .. 4*pg+3 of a block
    const int n_iter = C / 64;                   // 8 octets per pass
    for (int blk = 0; blk < CIN_NPB; ++blk) {
        const int l0 = l00 + blk * TP;
        const int tile = blockIdx.x * CIN_NPB + blk;
        if (l0 >= L) break;
        for (int it = 0; it < n_iter; ++it) {
            const int oct = it * 8 + (lane & 7);
            float acc[4][8];
            float cvv[MODE == 2 ? CIN_MAX_CC : 1][4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[u][j] = bs[oct * 8 + j];
            for (int ci = 0; ci < Cx; ++ci) {
                const float* xr = xs + ci * XP + blk * TP + pg * 4;
                const float4 xa4 = *reinterpret_cast<const float4*>(xr);
                const float2 xb2 = *reinterpret_cast<const float2*>(xr + 4);
                const float xv[6] = {xa4.x, xa4.y, xa4.z, xa4.w, xb2.x, xb2.y};
                if (MODE == 2) {
                    // input channels 1 .. Cc ARE the conditioning channels at full resolution (models.py:188-193 at level 0)
#pragma unroll
                    for (int j = 0; j < CIN_MAX_CC; ++j)
                        if (ci == 1 + j && j < ap.Cc) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) cvv[j][u] = xv[u + 1];
                        }
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float* wp = ws + (ci * 3 + k) * C + oct * 4;
                    const float4 wa = *reinterpret_cast<const float4*>(wp);
                    const float4 wb = *reinterpret_cast<const float4*>(wp + C / 2);
                    const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[u][j] = fmaf(xv[u + k], wv[j], acc[u][j]);
                }
            }
            if (MODE == 2) {
                T* outp = (T*)ap.out;
                T* poolp = (T*)ap.pooled;
                float cA[8], cB[8], cG[8], cE[8];
#pragma unroll
                for (int hsel = 0; hsel < 2; ++hsel) {
                    const float4 a4 = *reinterpret_cast<const float4*>(cf + 0 * C + hsel * (C / 2) + oct * 4);
                    const float4 b4 = *reinterpret_cast<const float4*>(cf + 1 * C + hsel * (C / 2) + oct * 4);
                    const float4 g4 = *reinterpret_cast<const float4*>(cf + 2 * C + hsel * (C / 2) + oct * 4);
                    const float4 e4 = *reinterpret_cast<const float4*>(cf + 3 * C + hsel * (C / 2) + oct * 4);
                    cA[hsel * 4 + 0] = a4.x; cA[hsel * 4 + 1] = a4.y; cA[hsel * 4 + 2] = a4.z; cA[hsel * 4 + 3] = a4.w;
                    cB[hsel * 4 + 0] = b4.x; cB[hsel * 4 + 1] = b4.y; cB[hsel * 4 + 2] = b4.z; cB[hsel * 4 + 3] = b4.w;
                    cG[hsel * 4 + 0] = g4.x; cG[hsel * 4 + 1] = g4.y; cG[hsel * 4 + 2] = g4.z; cG[hsel * 4 + 3] = g4.w;
                    cE[hsel * 4 + 0] = e4.x; cE[hsel * 4 + 1] = e4.y; cE[hsel * 4 + 2] = e4.z; cE[hsel * 4 + 3] = e4.w;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float z = fmaf(acc[u][j], cA[j], cB[j]);
                        const float sl = sizeof(T) == 2 ? silu_tanh(z) : silu_f<false>(z);
                        acc[u][j] = fmaf(sl, cG[j], cE[j]);
                    }
                for (int jc = 0; jc < ap.Cc; ++jc) {
                    const float4 w0 = *reinterpret_cast<const float4*>(cf + (4 + jc) * C + oct * 4);
                    const float4 w1 = *reinterpret_cast<const float4*>(cf + (4 + jc) * C + C / 2 + oct * 4);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                    float cu[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                    for (int j = 0; j < CIN_MAX_CC; ++j)
                        if (j == jc) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) cu[u] = cvv[j][u];
                        }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[u][j] = fmaf(wv[j], cu[u], acc[u][j]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int l = l0 + pg * 4 + u;
                    if (l < L) st8(outp + ((size_t)b * L + l) * C + oct * 8, acc[u]);
                }
                if (poolp != nullptr) {
                    const int Lp = L / 2;
#pragma unroll
                    for (int u = 0; u < 4; u += 2) {
                        const int l = l0 + pg * 4 + u;
                        if (l + 1 < L) {
                            float pv[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) pv[j] = 0.5f * (acc[u][j] + acc[u + 1][j]);
                            st8(poolp + ((size_t)b * Lp + (l >> 1)) * C + oct * 8, pv);
                        }
                    }
                }
                continue;
            }
            float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int l = l0 + pg * 4 + u;
                if (l < L) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (MODE == 0) acc[u][j] = round_to(acc[u][j], raw);
                        s1 += acc[u][j];
                        s2 += acc[u][j] * acc[u][j];
                    }
                    if (MODE == 0) st8(raw + ((size_t)b * L + l) * C + oct * 8, acc[u]);
                }
            }
            // fold the 4 position groups of the warp that share an octet (lanes differing in bits 3, 4)
            s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
            if (lane < 8) {
                wst[(oct * 8 + warp) * 2 + 0] = s1;
                wst[(oct * 8 + warp) * 2 + 1] = s2;
            }
        }
        if (MODE == 2) continue;
        __syncthreads();
        if (threadIdx.x < 8) {
            const int g = threadIdx.x, opg = C / 64;     // octets per GroupNorm group (C/8 channels per group)
            float a1 = 0.0f, a2 = 0.0f;
            for (int o = g * opg; o < (g + 1) * opg; ++o)
                for (int wi = 0; wi < 8; ++wi) {
                    a1 += wst[(o * 8 + wi) * 2 + 0];
                    a2 += wst[(o * 8 + wi) * 2 + 1];
                }
            float* pt = part + ((size_t)b * n_part + tile) * 16;
            pt[g * 2 + 0] = a1;
            pt[g * 2 + 1] = a2;
        }
        __syncthreads();
    }
}

// ---- training forward, bf16, C = 64, Cx <= 8: the same conv as warp-level tf32 MMAs (mma.sync.m16n8k8, fp32 accumulate).  With
// in_ch = 7 the CUDA-core kernel above is FMA-bound (1344 FMAs per row: 117 us at B = 256, L = 4096 against 21 us of HBM time).
// A warp owns 16 rows x 64 channels; K = 3*Cx padded to 8*KS.  MMA column n = 8 nt + c carries channel
//   ch(nt, c) = 32 (nt >> 2) + 8 (c >> 1) + 2 (nt & 3) + (c & 1),
// so the four accumulator pairs a lane holds for nt = 4h .. 4h+3 are 8 CONSECUTIVE channels (= one GroupNorm group, one 16-byte
// store) and the four lanes of a quad cover 64 contiguous bytes of a row.
__device__ __forceinline__ uint32_t cin_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__global__ void __launch_bounds__(256) conv_in_mma_kernel(const float* __restrict__ xa, const float* __restrict__ xb,
                                                          const int* __restrict__ step_ptr, int Cx, int L,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          bf16* __restrict__ raw, float* __restrict__ part, int n_part) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int C = 64, TP = 128, XP = CIN_NPB * TP + 8;
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                              // [Cx][XP]: xs[c][j] = x[c][l00 - 4 + j] (16-byte aligned body at j = 4)
    float* wsf = xs + Cx * XP;                   // [KS][8 n-tiles][32 lanes][2] tf32 B fragments
    const int KS = (3 * Cx + 7) / 8;
    float* bs = wsf + KS * 512;                  // [C]
    float* wst = bs + C;                         // [8 groups][8 warps][2]
    const int b = blockIdx.y, l00 = blockIdx.x * (CIN_NPB * TP);
    const int step = step_ptr != nullptr ? *step_ptr : 0;
    const float* x = (step & 1) ? xb : xa;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    // input staging: vector loads, all of a thread's loads in flight before the first store (the scalar div/mod loop of the
    // CUDA-core kernel cost more than the MMAs: profiles/r01j_ncu_full_in_mma.md)
    {
        constexpr int Q = CIN_NPB * TP / 4;      // float4 per channel
        constexpr int UNR = 4;
        for (int i0 = threadIdx.x; i0 < Cx * Q; i0 += 256 * UNR) {
            float4 v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int i = i0 + u * 256;
                const int c = i / Q, l = l00 + 4 * (i % Q);
                v[u] = (i < Cx * Q && l < L) ? *reinterpret_cast<const float4*>(x + ((size_t)b * Cx + c) * L + l)
                                             : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int i = i0 + u * 256;
                if (i < Cx * Q) *reinterpret_cast<float4*>(xs + (i / Q) * XP + 4 + 4 * (i % Q)) = v[u];
            }
        }
        if (threadIdx.x < 2 * Cx) {              // the two halo positions of every channel
            const int c = threadIdx.x >> 1, side = threadIdx.x & 1;
            const int l = side ? l00 + CIN_NPB * TP : l00 - 1;
            xs[c * XP + (side ? 4 + CIN_NPB * TP : 3)] = (l >= 0 && l < L) ? x[((size_t)b * Cx + c) * L + l] : 0.0f;
        }
    }
    for (int i = threadIdx.x; i < KS * 512; i += 256) {
        const int j = i & 1, ln = (i >> 1) & 31, nt = (i >> 6) & 7, ks = i >> 9;
        const int kk = ks * 8 + (ln & 3) + 4 * j, cc = ln >> 2;
        const int co = 32 * (nt >> 2) + 8 * (cc >> 1) + 2 * (nt & 3) + (cc & 1);
        wsf[i] = __uint_as_float(cin_tf32(kk < 3 * Cx ? w[(size_t)co * Cx * 3 + kk] : 0.0f));
    }
    for (int i = threadIdx.x; i < C; i += 256) bs[i] = bias[i];
    int aoff[3][2];
#pragma unroll
    for (int ks = 0; ks < 3; ++ks)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int kk = ks * 8 + t4 + 4 * j;
            aoff[ks][j] = kk < 3 * Cx ? (kk / 3) * XP + kk % 3 + 3 : -1;
        }
    __syncthreads();
    const float2* wf2 = reinterpret_cast<const float2*>(wsf);
    for (int blk = 0; blk < CIN_NPB; ++blk) {
        const int l0 = l00 + blk * TP;
        if (l0 >= L) break;
        const int tile = blockIdx.x * CIN_NPB + blk;
        const int rbase = blk * TP + warp * 16;          // this warp's 16 rows inside the CTA's staged input
        float acc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float2 b2 = *reinterpret_cast<const float2*>(bs + 32 * (nt >> 2) + 8 * t4 + 2 * (nt & 3));
            acc[nt][0] = b2.x; acc[nt][1] = b2.y; acc[nt][2] = b2.x; acc[nt][3] = b2.y;
        }
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
            if (ks < KS) {
                const float* x0 = xs + rbase + g;
                uint32_t af[4];
                af[0] = aoff[ks][0] >= 0 ? cin_tf32(x0[aoff[ks][0]]) : 0u;
                af[1] = aoff[ks][0] >= 0 ? cin_tf32(x0[aoff[ks][0] + 8]) : 0u;
                af[2] = aoff[ks][1] >= 0 ? cin_tf32(x0[aoff[ks][1]]) : 0u;
                af[3] = aoff[ks][1] >= 0 ? cin_tf32(x0[aoff[ks][1] + 8]) : 0u;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const float2 bf = wf2[(ks * 8 + nt) * 32 + lane];
                    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                                 : "r"(af[0]), "r"(af[1]), "r"(af[2]), "r"(af[3]), "r"(__float_as_uint(bf.x)), "r"(__float_as_uint(bf.y)));
                }
            }
        }
        // rows l0 + 16 warp + g (+8): channels 32 h + 8 t4 .. +7 (GroupNorm group 4 h + t4) as one 16-byte store per h
        float s1[2] = {0.0f, 0.0f}, s2[2] = {0.0f, 0.0f};
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int l = l0 + warp * 16 + g + 8 * rr;
            if (l < L) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t wd[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        wd[q] = pack_bf16x2(acc[4 * h + q][2 * rr], acc[4 * h + q][2 * rr + 1]);
                        const float lo = __uint_as_float(wd[q] << 16), hi = __uint_as_float(wd[q] & 0xffff0000u);
                        s1[h] += lo + hi;
                        s2[h] = fmaf(lo, lo, fmaf(hi, hi, s2[h]));
                    }
                    *reinterpret_cast<uint4*>(raw + ((size_t)b * L + l) * C + 32 * h + 8 * t4) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
                }
            }
        }
        // fold the 8 row lanes (g) of every (h, t4): lane bits 2..4
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                s1[h] += __shfl_xor_sync(0xffffffffu, s1[h], o);
                s2[h] += __shfl_xor_sync(0xffffffffu, s2[h], o);
            }
        }
        if (g == 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                wst[((4 * h + t4) * 8 + warp) * 2 + 0] = s1[h];
                wst[((4 * h + t4) * 8 + warp) * 2 + 1] = s2[h];
            }
        }
        __syncthreads();
        if (threadIdx.x < 8) {
            const int grp = threadIdx.x;
            float a1 = 0.0f, a2 = 0.0f;
            for (int wi = 0; wi < 8; ++wi) {
                a1 += wst[(grp * 8 + wi) * 2 + 0];
                a2 += wst[(grp * 8 + wi) * 2 + 1];
            }
            float* pt = part + ((size_t)b * n_part + tile) * 16;
            pt[grp * 2 + 0] = a1;
            pt[grp * 2 + 1] = a2;
        }
        __syncthreads();
    }
}

int g_conv_in_mma = 1;

template <typename T, int MODE>
static int conv_in_launch(const float* x, const float* x_alt, const int* step_ptr, int B, int Cx, int L, const float* w,
                          const float* bias, int C, void* raw, float* part, const ConvInApply& ap, cudaStream_t st) {
    constexpr int TP = 128, XP = CIN_NPB * TP + 8;
    const int n_part = gw_cdiv(L, TP);
    size_t smem = (size_t)(Cx * XP + Cx * 3 * C + C + (C / 8) * 8 * 2 + (MODE == 2 ? (4 + CIN_MAX_CC) * C : 0)) * sizeof(float);
    dim3 grid(gw_cdiv(L, CIN_NPB * TP), B);
    GW_CUDA(cudaFuncSetAttribute(conv_in_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GW_CUDA(gw_launch_pdl(conv_in_kernel<T, MODE>, grid, dim3(256), (size_t)(smem), st, x, x_alt ? x_alt : x, step_ptr, Cx, L, w, bias, C, (T*)raw, part, n_part, ap));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

extern "C" int gw_conv_in(const float* x, const float* x_alt, const int* step_ptr, int B, int Cx, int L, const float* w,
                          const float* bias, int C, void* raw, int dtype, float* part, void* stream) {
    GW_REQUIRE(C % 64 == 0 && C <= 256, "gw_conv_in: C=%d must be a multiple of 64 and <= 256", C);
    GW_REQUIRE(Cx >= 1 && Cx <= 16, "gw_conv_in: Cx=%d", Cx);
    GW_REQUIRE(dtype == GW_F32 || dtype == GW_BF16, "gw_conv_in: dtype %d", dtype);
    if (dtype == GW_BF16 && C == 64 && Cx <= 8 && L % 4 == 0 && g_conv_in_mma) {
        constexpr int TP = 128, XP = CIN_NPB * TP + 8;
        const int KS = (3 * Cx + 7) / 8, n_part = gw_cdiv(L, TP);
        const size_t smem = (size_t)(Cx * XP + KS * 512 + C + 128) * sizeof(float);
        dim3 grid(gw_cdiv(L, CIN_NPB * TP), B);
        GW_CUDA(cudaFuncSetAttribute(conv_in_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GW_CUDA(gw_launch_pdl(conv_in_mma_kernel, grid, dim3(256), (size_t)(smem), (cudaStream_t)stream, x, x_alt ? x_alt : x, step_ptr, Cx, L, w, bias, (bf16*)raw, part,
                                                                     n_part));
        GW_LAUNCH_CHECK();
        return GW_OK;
    }
    ConvInApply ap;
    memset(&ap, 0, sizeof(ap));
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GW_F32) return conv_in_launch<float, 0>(x, x_alt, step_ptr, B, Cx, L, w, bias, C, raw, part, ap, st);
    return conv_in_launch<bf16, 0>(x, x_alt, step_ptr, B, Cx, L, w, bias, C, raw, part, ap, st);
}

// Fused first block for inference (no raw tensor): pass 1 = GroupNorm partial sums of the conv output, pass 2 = recompute
// the conv and apply GroupNorm + SiLU + cond 1x1 conv + FiLM (+ avg_pool) -> out [B, L, C], pooled [B, L/2, C] or NULL.
// part: scratch fp32 [B, ceil(L/128), 8, 2].  The conditioning channels are x[:, 1:1+Cc] themselves (level-0 interpolation
// is the identity), so no conditioning pyramid is read.
extern "C" int gw_conv_in_block(const float* x, const float* x_alt, const int* step_ptr, int B, int Cx, int L, const float* w,
                                const float* bias, int C, const float* gn_w, const float* gn_b, int Cc, const float* wc,
                                const float* bc, const float* film, int film_off, long film_b_stride, long film_step_stride,
                                void* out, void* pooled, int dtype, float* part, void* stream) {
    GW_REQUIRE(C % 64 == 0 && C <= 256, "gw_conv_in_block: C=%d must be a multiple of 64 and <= 256", C);
    GW_REQUIRE(Cx >= 1 && Cx <= 16 && Cc >= 0 && Cc <= CIN_MAX_CC && 1 + Cc <= Cx, "gw_conv_in_block: Cx=%d Cc=%d", Cx, Cc);
    GW_REQUIRE(dtype == GW_F32 || dtype == GW_BF16, "gw_conv_in_block: dtype %d", dtype);
    ConvInApply ap;
    memset(&ap, 0, sizeof(ap));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = dtype == GW_F32 ? conv_in_launch<float, 1>(x, x_alt, step_ptr, B, Cx, L, w, bias, C, nullptr, part, ap, st)
                             : conv_in_launch<bf16, 1>(x, x_alt, step_ptr, B, Cx, L, w, bias, C, nullptr, part, ap, st);
    if (rc != GW_OK) return rc;
    ap.part_in = part; ap.gn_w = gn_w; ap.gn_b = gn_b; ap.wc = wc; ap.bc = bc; ap.film = film; ap.film_off = film_off;
    ap.film_b_stride = film_b_stride; ap.film_step_stride = film_step_stride; ap.Cc = Cc; ap.out = out; ap.pooled = pooled;
    return dtype == GW_F32 ? conv_in_launch<float, 2>(x, x_alt, step_ptr, B, Cx, L, w, bias, C, nullptr, nullptr, ap, st)
                           : conv_in_launch<bf16, 2>(x, x_alt, step_ptr, B, Cx, L, w, bias, C, nullptr, nullptr, ap, st);
}

// ------------------------------------------------------------------------------------------------
// exact-mode SIMT conv (K2-K7 in fp32): 64 positions x 64 couts per CTA, K chunks of 16 channels x 3 taps.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv3_simt_kernel(const T* __restrict__ src0, int C0, int L0, int up0,
                                                         const T* __restrict__ src1, int C1, int L,
                                                         const float* __restrict__ w3, const float* __restrict__ bias,
                                                         int Cout, T* __restrict__ raw, float* __restrict__ part,
                                                         int n_part) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int TP = 64, TN = 64, KC = 16;
    __shared__ __align__(16) float xs[(TP + 2) * KC];
    __shared__ __align__(16) float ws[3 * KC * TN];
    __shared__ float red[256 * 3];
    const int b = blockIdx.z, n0 = blockIdx.y * TN, tile = blockIdx.x, l0 = tile * TP;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int Cin = C0 + C1;
    float acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[p][n] = 0.0f;

    for (int c0 = 0; c0 < Cin; c0 += KC) {
        // input tile: rows l0-1 .. l0+TP, channels c0..c0+15 (two octets per row)
        for (int i = threadIdx.x; i < (TP + 2) * 2; i += blockDim.x) {
            const int row = i >> 1, oct = i & 1;
            const int l = l0 + row - 1;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.0f;
            if (l >= 0 && l < L) {
                if (c0 < C0) {
                    const int ls = up0 ? (l >> 1) : l;
                    if (ls < L0) ld8(src0 + ((size_t)b * L0 + ls) * C0 + c0 + oct * 8, v);
                } else {
                    ld8(src1 + ((size_t)b * L + l) * C1 + (c0 - C0) + oct * 8, v);
                }
            }
            float* d = xs + row * KC + oct * 8;
            *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        // weights: ws[k][ci][n] = w[n0+n][c0+ci][k]  (reference layout [Cout, Cin, 3]: 48 contiguous floats per n)
        for (int i = threadIdx.x; i < 3 * KC * TN; i += blockDim.x) {
            const int n = i / (3 * KC), j = i % (3 * KC);
            const int ci = j / 3, k = j - ci * 3;
            ws[(k * KC + ci) * TN + n] = w3[((size_t)(n0 + n) * Cin + c0 + ci) * 3 + k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll 4
            for (int ci = 0; ci < KC; ++ci) {
                const float4 wv = *reinterpret_cast<const float4*>(ws + (k * KC + ci) * TN + tx * 4);
                float a[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) a[p] = xs[(ty * 4 + p + k) * KC + ci];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    acc[p][0] = fmaf(a[p], wv.x, acc[p][0]);
                    acc[p][1] = fmaf(a[p], wv.y, acc[p][1]);
                    acc[p][2] = fmaf(a[p], wv.z, acc[p][2]);
                    acc[p][3] = fmaf(a[p], wv.w, acc[p][3]);
                }
            }
        }
        __syncthreads();
    }
    float bb[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (bias != nullptr) {
        const float4 bv = *reinterpret_cast<const float4*>(bias + n0 + tx * 4);
        bb[0] = bv.x; bb[1] = bv.y; bb[2] = bv.z; bb[3] = bv.w;
    }
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int l = l0 + ty * 4 + p;
        if (l < L) {
            float o[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                o[n] = round_to(acc[p][n] + bb[n], raw);
                s1 += o[n];
                s2 += o[n] * o[n];
            }
            T* dst = raw + ((size_t)b * L + l) * Cout + n0 + tx * 4;
            if (sizeof(T) == 4) {
                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
                uint2 pk;
                pk.x = pack_bf16x2(o[0], o[1]);
                pk.y = pack_bf16x2(o[2], o[3]);
                *reinterpret_cast<uint2*>(dst) = pk;
            }
        }
    }
    if (part == nullptr) return;                      // dgrad use: no GroupNorm statistics wanted
    const int cg = Cout / 8;
    const int n_groups = TN / cg > 0 ? TN / cg : 1;   // groups covered by this cout tile (cg <= 64)
    tile_stats_reduce(s1, s2, (tx * 4) / cg, n_groups, red, part + ((size_t)b * n_part + tile) * 16, n0 / cg);
}

extern "C" int gw_conv3_simt(const void* src0, int C0, int L0, int up0, const void* src1, int C1, int B, int L,
                             const float* w3, const float* bias, int Cout, void* raw, int dtype, float* part,
                             void* stream) {
    GW_REQUIRE(C0 % 16 == 0 && C1 % 16 == 0 && C0 > 0, "gw_conv3_simt: channels C0=%d C1=%d", C0, C1);
    GW_REQUIRE(Cout % 64 == 0 && Cout <= 4096, "gw_conv3_simt: Cout=%d", Cout);
    GW_REQUIRE((src1 != nullptr) == (C1 > 0), "gw_conv3_simt: src1/C1 mismatch");
    GW_REQUIRE(dtype == GW_F32 || dtype == GW_BF16, "gw_conv3_simt: dtype %d", dtype);
    const int n_part = gw_cdiv(L, 64);
    dim3 grid(n_part, Cout / 64, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GW_F32)
        GW_CUDA(gw_launch_pdl(conv3_simt_kernel<float>, grid, dim3(256), (size_t)(0), st, (const float*)src0, C0, L0, up0, (const float*)src1, C1, L, w3, bias,
                                                        Cout, (float*)raw, part, n_part));
    else
        GW_CUDA(gw_launch_pdl(conv3_simt_kernel<bf16>, grid, dim3(256), (size_t)(0), st, (const bf16*)src0, C0, L0, up0, (const bf16*)src1, C1, L, w3, bias,
                                                       Cout, (bf16*)raw, part, n_part));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------
// fused GroupNorm-apply + SiLU + cond 1x1 conv + FiLM (+ pooled output)
// ------------------------------------------------------------------------------------------------
#define GN_MAX_CC 8
// 4-channel vectors: 16 B (fp32) or 8 B (bf16) per lane, a warp always touches whole 128 B lines
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void ld4(const bf16* p, float (&v)[4]) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(bf16* p, const float (&v)[4]) {
    uint2 r;
    r.x = pack_bf16x2(v[0], v[1]);
    r.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = r;
}

// CC = compile-time number of conditioning channels (0, 1, 5) or -1 for the generic (<= GN_MAX_CC) path.
// A thread owns one channel quad for the whole CTA (coefficients live in registers) and walks row PAIRS so the
// avg_pool1d(2,2) output is formed from fp32 values; UN pairs are in flight per iteration to cover HBM latency.
template <typename T, bool FAST, int CC>
__global__ void __launch_bounds__(256, FAST ? ((CC == 0 || CC == 1) ? 3 : 2) : 2)
gn_apply_kernel(const T* __restrict__ raw, const float* __restrict__ part, int n_part, int L, int C,
                const float* __restrict__ gn_w, const float* __restrict__ gn_b, const float* __restrict__ cond, int Cc_rt,
                const float* __restrict__ wc, const float* __restrict__ bc, const float* __restrict__ film, int film_off,
                long film_b_stride, long film_step_stride, const int* __restrict__ step_ptr, T* __restrict__ out,
                T* __restrict__ pooled, float* __restrict__ stats_out, int rows_per_cta) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int NC = CC >= 0 ? CC : GN_MAX_CC;
    constexpr int NCA = NC > 0 ? NC : 1;
    constexpr int UN = (FAST && (CC == 0 || CC == 1)) ? 4 : 2;
    const int Cc = CC >= 0 ? CC : Cc_rt;
    __shared__ float s_mean[8], s_rstd[8];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cg = C / 8;
    if (warp < 8) {
        double a1 = 0.0, a2 = 0.0;
        const float* pp = part + (size_t)b * n_part * 16 + warp * 2;
        for (int i = lane; i < n_part; i += 32) {
            a1 += (double)pp[(size_t)i * 16];
            a2 += (double)pp[(size_t)i * 16 + 1];
        }
        a1 = warp_sum_d(a1);
        a2 = warp_sum_d(a2);
        if (lane == 0) {
            const double n = (double)cg * (double)L;
            const double mean = a1 / n;
            double var = a2 / n - mean * mean;
            if (var < 0.0) var = 0.0;
            const float rstd = (float)(1.0 / sqrt(var + 1e-5));
            s_mean[warp] = (float)mean;
            s_rstd[warp] = rstd;
            if (stats_out != nullptr && blockIdx.x == 0) {
                stats_out[((size_t)b * 8 + warp) * 2 + 0] = (float)mean;
                stats_out[((size_t)b * 8 + warp) * 2 + 1] = rstd;
            }
        }
    }
    __syncthreads();
    const int n_quad = C / 4;                   // 16..512; host guarantees it divides or is a multiple of 256
    const int step = step_ptr != nullptr ? *step_ptr : 0;
    const float* fr = film + (size_t)step * film_step_stride + (size_t)b * film_b_stride + film_off;
    const int r0 = blockIdx.x * rows_per_cta;
    const bool do_pool = pooled != nullptr;
    const int Lp = L / 2;
    const int n_pairs = rows_per_cta / 2;
    const int pr_stride = n_quad >= 256 ? 1 : 256 / n_quad;
    for (int quad = threadIdx.x % n_quad; quad < n_quad; quad += 256) {
        const int pr0 = n_quad >= 256 ? 0 : threadIdx.x / n_quad;
        // h = silu(A*x + Bn) + (bc + sum_j wc[j]*cond[j]);  out = h*G + E   (models.py:165-166, 205, 173)
        // FAST folds the second line into the first: out = silu(.)*G + (E + bc*G) + sum_j (wc[j]*G)*cond[j]
        float cA[4], cB[4], cG[4], cE[4], cC[4], cW[4][NCA];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = quad * 4 + i;
            const int g = c / cg;
            const float a = s_rstd[g] * gn_w[c];
            cA[i] = a;
            cB[i] = gn_b[c] - s_mean[g] * a;
            cG[i] = 1.0f + fr[c];
            cE[i] = fr[C + c];
            cC[i] = NC > 0 ? bc[c] : 0.0f;
#pragma unroll
            for (int j = 0; j < NCA; ++j) cW[i][j] = (NC > 0 && j < Cc) ? wc[c * Cc + j] : 0.0f;
            if (FAST) {
                cE[i] = fmaf(cC[i], cG[i], cE[i]);
#pragma unroll
                for (int j = 0; j < NCA; ++j) cW[i][j] *= cG[i];
            }
        }
        const T* rbase = raw + (size_t)b * L * C + quad * 4;
        T* obase = out + (size_t)b * L * C + quad * 4;
        const float* cbase = cond + (size_t)b * L * Cc;
        for (int pr = pr0; pr < n_pairs; pr += pr_stride * UN) {
            float x[UN][2][4];
            float cv[UN][2][NCA];
            // all loads first (addresses clamped into the sample so they need no predicate)
#pragma unroll
            for (int u = 0; u < UN; ++u) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int l = r0 + 2 * (pr + u * pr_stride) + h;
                    l = l < L ? l : L - 1;
                    ld4(rbase + (size_t)l * C, x[u][h]);
                    if (NC > 0) {
#pragma unroll
                        for (int j = 0; j < NCA; ++j) cv[u][h][j] = (j < Cc) ? cbase[(size_t)l * Cc + j] : 0.0f;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int l0 = r0 + 2 * (pr + u * pr_stride);
                const bool pair_ok = (pr + u * pr_stride) < n_pairs;
                float o[2][4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float z = fmaf(x[u][h][i], cA[i], cB[i]);
                        if (FAST) {
                            float v = fmaf(silu_tanh(z), cG[i], cE[i]);
                            if (NC > 0) {
#pragma unroll
                                for (int j = 0; j < NCA; ++j) v = fmaf(cW[i][j], cv[u][h][j], v);
                            }
                            o[h][i] = v;
                        } else {
                            float v = silu_f<false>(z);
                            if (NC > 0) {
                                float cb = cC[i];
#pragma unroll
                                for (int j = 0; j < NCA; ++j) cb = fmaf(cW[i][j], cv[u][h][j], cb);
                                v += cb;
                            }
                            o[h][i] = fmaf(v, cG[i], cE[i]);
                        }
                    }
                    if (pair_ok && l0 + h < L) st4(obase + (size_t)(l0 + h) * C, o[h]);
                }
                if (do_pool && pair_ok && l0 + 1 < L) {
                    float pv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) pv[i] = 0.5f * (o[0][i] + o[1][i]);
                    st4(pooled + ((size_t)b * Lp + (l0 >> 1)) * C + quad * 4, pv);
                }
            }
        }
    }
}

// bf16 fast path of the kernel above on packed fp32x2 math (the scalar version is issue-bound at ~60 % of HBM speed):
//   out = silu(A x + Bn) G + (E + bc G) + sum_j (wc_j G) cond_j,   silu(z) = h + h tanh(h), h = z/2 (one MUFU per element)
template <int CC>
__global__ void __launch_bounds__(256, (CC == 0 || CC == 1) ? 3 : 2)
gn_apply_bf16_kernel(const bf16* __restrict__ raw, const float* __restrict__ part, int n_part, int L, int C,
                     const float* __restrict__ gn_w, const float* __restrict__ gn_b, const float* __restrict__ cond, int Cc_rt,
                     const float* __restrict__ wc, const float* __restrict__ bc, const float* __restrict__ film, int film_off,
                     long film_b_stride, long film_step_stride, const int* __restrict__ step_ptr, bf16* __restrict__ out,
                     bf16* __restrict__ pooled, float* __restrict__ stats_out, int rows_per_cta) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int NC = CC >= 0 ? CC : GN_MAX_CC;
    constexpr int NCA = NC > 0 ? NC : 1;
    constexpr int UN = (CC == 0 || CC == 1) ? 4 : 2;
    const int Cc = CC >= 0 ? CC : Cc_rt;
    __shared__ float s_mean[8], s_rstd[8];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cg = C / 8;
    if (warp < 8) {
        double a1 = 0.0, a2 = 0.0;
        const float* pp = part + (size_t)b * n_part * 16 + warp * 2;
        for (int i = lane; i < n_part; i += 32) {
            a1 += (double)pp[(size_t)i * 16];
            a2 += (double)pp[(size_t)i * 16 + 1];
        }
        a1 = warp_sum_d(a1);
        a2 = warp_sum_d(a2);
        if (lane == 0) {
            const double n = (double)cg * (double)L;
            const double mean = a1 / n;
            double var = a2 / n - mean * mean;
            if (var < 0.0) var = 0.0;
            const float rstd = (float)(1.0 / sqrt(var + 1e-5));
            s_mean[warp] = (float)mean;
            s_rstd[warp] = rstd;
            if (stats_out != nullptr && blockIdx.x == 0) {
                stats_out[((size_t)b * 8 + warp) * 2 + 0] = (float)mean;
                stats_out[((size_t)b * 8 + warp) * 2 + 1] = rstd;
            }
        }
    }
    __syncthreads();
    const int n_quad = C / 4;
    const int step = step_ptr != nullptr ? *step_ptr : 0;
    const float* fr = film + (size_t)step * film_step_stride + (size_t)b * film_b_stride + film_off;
    const int r0 = blockIdx.x * rows_per_cta;
    const bool do_pool = pooled != nullptr;
    const int Lp = L / 2;
    const int n_pairs = rows_per_cta / 2;
    const int pr_stride = n_quad >= 256 ? 1 : 256 / n_quad;
    const f32x2 half2 = pkf2(0.5f, 0.5f);
    for (int quad = threadIdx.x % n_quad; quad < n_quad; quad += 256) {
        const int pr0 = n_quad >= 256 ? 0 : threadIdx.x / n_quad;
        f32x2 hA[2], hB[2], G[2], E[2], W[NCA][2];
        {
            float a_[4], b_[4], g_[4], e_[4], w_[NCA][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = quad * 4 + i;
                const int g = c / cg;
                const float a = s_rstd[g] * gn_w[c];
                a_[i] = 0.5f * a;
                b_[i] = 0.5f * (gn_b[c] - s_mean[g] * a);
                g_[i] = 1.0f + fr[c];
                e_[i] = fmaf(NC > 0 ? bc[c] : 0.0f, g_[i], fr[C + c]);
#pragma unroll
                for (int j = 0; j < NCA; ++j) w_[j][i] = (NC > 0 && j < Cc) ? wc[c * Cc + j] * g_[i] : 0.0f;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                hA[h] = pkf2(a_[2 * h], a_[2 * h + 1]);
                hB[h] = pkf2(b_[2 * h], b_[2 * h + 1]);
                G[h] = pkf2(g_[2 * h], g_[2 * h + 1]);
                E[h] = pkf2(e_[2 * h], e_[2 * h + 1]);
#pragma unroll
                for (int j = 0; j < NCA; ++j) W[j][h] = pkf2(w_[j][2 * h], w_[j][2 * h + 1]);
            }
        }
        const bf16* rbase = raw + (size_t)b * L * C + quad * 4;
        bf16* obase = out + (size_t)b * L * C + quad * 4;
        const float* cbase = cond + (size_t)b * L * Cc;
        for (int pr = pr0; pr < n_pairs; pr += pr_stride * UN) {
            uint2 x[UN][2];
            float cv[UN][2][NCA];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int l = r0 + 2 * (pr + u * pr_stride) + h;
                    l = l < L ? l : L - 1;
                    x[u][h] = *reinterpret_cast<const uint2*>(rbase + (size_t)l * C);
                    if (NC > 0) {
#pragma unroll
                        for (int j = 0; j < NCA; ++j) cv[u][h][j] = (j < Cc) ? cbase[(size_t)l * Cc + j] : 0.0f;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int l0 = r0 + 2 * (pr + u * pr_stride);
                const bool pair_ok = (pr + u * pr_stride) < n_pairs;
                f32x2 o[2][2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t w2[2] = {x[u][h].x, x[u][h].y};
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const f32x2 xv = pk2(w2[q] << 16, w2[q] & 0xffff0000u);
                        const f32x2 hh = ffma2(xv, hA[q], hB[q]);
                        float h0, h1, t0, t1;
                        upk2(hh, h0, h1);
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
                        f32x2 v = ffma2(ffma2(hh, pkf2(t0, t1), hh), G[q], E[q]);
                        if (NC > 0) {
#pragma unroll
                            for (int j = 0; j < NCA; ++j) v = ffma2(W[j][q], pkf2(cv[u][h][j], cv[u][h][j]), v);
                        }
                        o[h][q] = v;
                    }
                    if (pair_ok && l0 + h < L) {
                        float a0, a1, a2, a3;
                        upk2(o[h][0], a0, a1);
                        upk2(o[h][1], a2, a3);
                        uint2 r;
                        r.x = pack_bf16x2(a0, a1);
                        r.y = pack_bf16x2(a2, a3);
                        *reinterpret_cast<uint2*>(obase + (size_t)(l0 + h) * C) = r;
                    }
                }
                if (do_pool && pair_ok && l0 + 1 < L) {
                    float a0, a1, a2, a3;
                    upk2(fmul2(fadd2(o[0][0], o[1][0]), half2), a0, a1);
                    upk2(fmul2(fadd2(o[0][1], o[1][1]), half2), a2, a3);
                    uint2 r;
                    r.x = pack_bf16x2(a0, a1);
                    r.y = pack_bf16x2(a2, a3);
                    *reinterpret_cast<uint2*>(pooled + ((size_t)b * Lp + (l0 >> 1)) * C + quad * 4) = r;
                }
            }
        }
    }
}

template <typename T, bool FAST>
static int gn_apply_launch(dim3 grid, cudaStream_t st, const void* raw, const float* part, int n_part, int L, int C,
                           const float* gn_w, const float* gn_b, const float* cond, int Cc, const float* wc,
                           const float* bc, const float* film, int film_off, long fbs, long fss, const int* step_ptr,
                           void* out, void* pooled, float* stats_out, int rows) {
#define GN_GO(CCV)                                                                                                     \
    GW_CUDA(gw_launch_pdl(gn_apply_kernel<T, FAST, CCV>, grid, dim3(256), (size_t)(0), st, (const T*)raw, part, n_part, L, C, gn_w, gn_b, cond, Cc, wc, bc, \
                                                        film, film_off, fbs, fss, step_ptr, (T*)out, (T*)pooled,        \
                                                        stats_out, rows))
    if (Cc == 0) GN_GO(0);
    else if (Cc == 1) GN_GO(1);
    else if (Cc == 5) GN_GO(5);
    else GN_GO(-1);
#undef GN_GO
    GW_LAUNCH_CHECK();
    return GW_OK;
}

extern "C" int gw_gn_apply(const void* raw, const float* part, int n_part, int B, int L, int C, const float* gn_w,
                           const float* gn_b, const float* cond, int Cc, const float* wc, const float* bc,
                           const float* film, int film_off, long film_b_stride, long film_step_stride,
                           const int* step_ptr, void* out, void* pooled, float* stats_out, int dtype, void* stream) {
    GW_REQUIRE(C % 64 == 0 && C <= 4096 && (256 % (C / 4) == 0 || (C / 4) % 256 == 0), "gw_gn_apply: C=%d", C);
    GW_REQUIRE(Cc >= 0 && Cc <= GN_MAX_CC, "gw_gn_apply: Cc=%d (max %d)", Cc, GN_MAX_CC);
    GW_REQUIRE((cond != nullptr) == (Cc > 0), "gw_gn_apply: cond/Cc mismatch");
    GW_REQUIRE(dtype == GW_F32 || dtype == GW_BF16, "gw_gn_apply: dtype %d", dtype);
    // >= 8 row pairs per thread so the per-thread coefficient setup is amortised
    int rows = C / 4 >= 256 ? 32 : 2 * 16 * (256 / (C / 4));     // 16 row pairs per thread
    if (rows > L) rows = (L + 1) & ~1;
    if (rows < 2) rows = 2;
    dim3 grid(gw_cdiv(L, rows), B);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GW_F32)
        return gn_apply_launch<float, false>(grid, st, raw, part, n_part, L, C, gn_w, gn_b, cond, Cc, wc, bc, film, film_off,
                                             film_b_stride, film_step_stride, step_ptr, out, pooled, stats_out, rows);
#define GNB_GO(CCV)                                                                                                         \
    GW_CUDA(gw_launch_pdl(gn_apply_bf16_kernel<CCV>, grid, dim3(256), (size_t)(0), st, (const bf16*)raw, part, n_part, L, C, gn_w, gn_b, cond, Cc, wc, bc, film,   \
                                                    film_off, film_b_stride, film_step_stride, step_ptr, (bf16*)out,            \
                                                    (bf16*)pooled, stats_out, rows))
    if (Cc == 0) GNB_GO(0);
    else if (Cc == 1) GNB_GO(1);
    else if (Cc == 5) GNB_GO(5);
    else GNB_GO(-1);
#undef GNB_GO
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------
// head conv (C+1 -> 1, k=3) fused with CFG combine + DDIM/DDPM update
// ------------------------------------------------------------------------------------------------
struct StepArgs {
    int mode, cfg_both, selfcond, pred_x0;
    float eps_scale, dc_weight;
    const float* y_dc;
    unsigned long long seed;
    long sample0;
    const unsigned long long* rng;      // device {seed, sample0} (graph-replay safe) or NULL -> the by-value fields
};

// CTA = 256 rows of h (254 output positions + the two halo rows), 256 threads.  Phase 1: every thread first issues the
// loads its epilogue will need (x_t taps, y, noise), then streams its share of h rows (4 x 16 B in flight, twice) and
// leaves the three per-row tap dots in smem.  Phase 2: thread i finishes position i (CFG combine, DDIM/DDPM update).
#define FS_ROWS 256
#define FS_TP (FS_ROWS - 2)
template <typename T>
__global__ void __launch_bounds__(256) final_step_kernel(const T* __restrict__ h, const float* __restrict__ net_a,
                                                         const float* __restrict__ net_b, int B, int Cx, int L, int C,
                                                         const float* __restrict__ wf, const float* __restrict__ bf,
                                                         StepArgs p, const float* __restrict__ coef,
                                                         const int* __restrict__ step_ptr, const float* __restrict__ noise,
                                                         float* __restrict__ eps_out, float* __restrict__ x0_out) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float pd[2][3][FS_ROWS];           // [half][tap][row] partial dots
    const int b = blockIdx.y, l0 = blockIdx.x * FS_TP;
    const int step = step_ptr != nullptr ? *step_ptr : 0;
    const float* net_in = (step & 1) ? net_b : net_a;
    float* net_out = const_cast<float*>((step & 1) ? net_a : net_b);
    const int n_half = (p.mode == 1 && p.cfg_both) ? 2 : 1;
    // ---- epilogue operands, requested before the h stream so their latency is hidden
    const int l = l0 + threadIdx.x;
    const bool mine = threadIdx.x < FS_TP && l < L;
    float xin[2][3] = {{0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}};
    float ydc = 0.0f, zin = 0.0f;
    const float* cf = coef != nullptr ? coef + (size_t)step * 16 : nullptr;
    float cfv[10] = {0.0f, 1.0f, 1.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f};
    if (mine) {
        for (int hf = 0; hf < n_half; ++hf) {
            const float* xr = net_in + (size_t)(b + hf * B) * Cx * L;     // channel 0 = x_t
            xin[hf][0] = l > 0 ? xr[l - 1] : 0.0f;
            xin[hf][1] = xr[l];
            xin[hf][2] = l + 1 < L ? xr[l + 1] : 0.0f;
        }
        if (p.mode == 1) {
#pragma unroll
            for (int i = 0; i < 10; ++i) cfv[i] = cf[i];
            if (p.dc_weight > 0.0f) ydc = p.y_dc[(size_t)b * L + l];
            if (noise != nullptr && cfv[4] > 0.0f && cfv[7] == 0.0f) zin = noise[((size_t)(int)cfv[8] * B + b) * L + l];
        }
    }
    // ---- phase 1: tap dots of rows l0-1 .. l0+254
    const int n_oct = C / 8;                 // threads per row
    const int rpp = 256 / n_oct;             // rows per pass
    const int oct = threadIdx.x % n_oct, tr = threadIdx.x / n_oct;
    float w0[8], w1[8], w2[8];               // this thread's channel octet of the three taps
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = oct * 8 + i;
        w0[i] = wf[c * 3 + 0];
        w1[i] = wf[c * 3 + 1];
        w2[i] = wf[c * 3 + 2];
    }
    for (int hf = 0; hf < n_half; ++hf) {
        const T* hb = h + (size_t)(b + hf * B) * L * C + oct * 8;
        for (int r0 = 0; r0 < FS_ROWS; r0 += rpp * 4) {
            float v[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int lr = l0 + r0 + u * rpp + tr - 1;
                lr = lr < 0 ? 0 : (lr >= L ? L - 1 : lr);
                ld8(hb + (size_t)lr * C, v[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + u * rpp + tr;
                const int lr = l0 + r - 1;
                const bool ok = r < FS_ROWS && lr >= 0 && lr < L;
                float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    d0 = fmaf(v[u][i], w0[i], d0);
                    d1 = fmaf(v[u][i], w1[i], d1);
                    d2 = fmaf(v[u][i], w2[i], d2);
                }
                if (!ok) { d0 = 0.0f; d1 = 0.0f; d2 = 0.0f; }
                for (int o = n_oct >> 1; o > 0; o >>= 1) {
                    d0 += __shfl_xor_sync(0xffffffffu, d0, o);
                    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
                    d2 += __shfl_xor_sync(0xffffffffu, d2, o);
                }
                if (oct == 0 && r < FS_ROWS) {
                    pd[hf][0][r] = d0;
                    pd[hf][1][r] = d1;
                    pd[hf][2][r] = d2;
                }
            }
        }
    }
    __syncthreads();
    if (!mine) return;
    // ---- phase 2
    const float wx0 = wf[C * 3 + 0], wx1 = wf[C * 3 + 1], wx2 = wf[C * 3 + 2], bias = bf[0];
    float outv[2] = {0.0f, 0.0f};
    const float xt_c = xin[0][1];
    for (int hf = 0; hf < n_half; ++hf) {
        const int r = threadIdx.x + 1;
        float acc = pd[hf][0][r - 1] + pd[hf][1][r] + pd[hf][2][r + 1];
        acc += fmaf(xin[hf][0], wx0, fmaf(xin[hf][1], wx1, xin[hf][2] * wx2));
        outv[hf] = acc + bias;
    }
    if (p.mode == 0) {
        eps_out[(size_t)b * L + l] = outv[0];
        return;
    }
    const float c_s1mab = cfv[0], c_sab = cfv[1], c_sabp = cfv[2], c_dir = cfv[3], c_sig = cfv[4], c_w = cfv[5];
    const int use = (int)cfv[6], last = (int)cfv[7];
    const float c_s1mab_cl = cfv[9];
    float o;
    if (use == 0) o = outv[0];
    else if (use == 1) o = p.cfg_both ? outv[1] : outv[0];
    else o = __fadd_rn(outv[1], __fmul_rn(c_w, __fsub_rn(outv[0], outv[1])));       // inference.py:460
    float eps, x0;
    if (!p.pred_x0) {
        eps = __fmul_rn(p.eps_scale, o);
        x0 = __fdiv_rn(__fsub_rn(xt_c, __fmul_rn(c_s1mab, eps)), c_sab);          // inference.py:465-466 (no FMA contraction)
    } else {
        x0 = o;
        eps = __fdiv_rn(__fsub_rn(xt_c, __fmul_rn(c_sab, x0)), c_s1mab_cl);       // inference.py:468-469
    }
    if (p.dc_weight > 0.0f)
        x0 = __fadd_rn(__fmul_rn(1.0f - p.dc_weight, x0), __fmul_rn(p.dc_weight, ydc));   // inference.py:472
    float xn;
    if (last) {
        xn = x0;
    } else {
        float nz = 0.0f;
        if (c_sig > 0.0f) {
            float z = zin;
            if (noise == nullptr) {
                float z4[4];
                const unsigned long long sd = p.rng != nullptr ? p.rng[0] : p.seed;
                const long s0 = p.rng != nullptr ? (long)p.rng[1] : p.sample0;
                Philox::normal4(sd, (uint32_t)(s0 + b), (uint32_t)step + 1u, (uint32_t)(l >> 2), z4);
                z = z4[l & 3];
            }
            nz = __fmul_rn(c_sig, z);
        }
        xn = __fadd_rn(__fadd_rn(__fmul_rn(c_sabp, x0), __fmul_rn(c_dir, eps)), nz);   // inference.py:481-484
    }
    for (int hf = 0; hf < n_half; ++hf) {
        float* orow = net_out + (size_t)(b + hf * B) * Cx * L;
        orow[l] = xn;
        if (p.selfcond) orow[(size_t)(Cx - 1) * L + l] = x0;
    }
    if (eps_out != nullptr) eps_out[(size_t)b * L + l] = eps;
    if (x0_out != nullptr) x0_out[(size_t)b * L + l] = x0;
}

int final_step_dots(const void* dots, const float* net_a, const float* net_b, int B, int Cx, int L, int C, const float* wf,
                    const float* bf, const gw_step_params* p, const float* coef, const int* step_ptr, const float* noise,
                    float* eps_out, float* x0_out, cudaStream_t st);
int final_step_stream(const void* h, const float* net_a, const float* net_b, int B, int Cx, int L, const float* wf, const float* bf,
                      const gw_step_params* p, const float* coef, const int* step_ptr, const float* noise, float* eps_out,
                      float* x0_out, cudaStream_t st);
int g_final_stream = 1;
int g_pdl = 0;      // flipped to 1 once verified on the GPU (see profiles)

extern "C" int gw_final_step(const void* h, int dtype, const float* net_a, const float* net_b, int B, int Cx, int L,
                             int C, const float* wf, const float* bf, const gw_step_params* p, const float* coef,
                             const int* step_ptr, const float* noise, float* eps_out, float* x0_out, void* stream) {
    GW_REQUIRE(p != nullptr, "gw_final_step: params");
    GW_REQUIRE(C % 64 == 0 && C <= 256 && ((C / 8) & (C / 8 - 1)) == 0, "gw_final_step: C=%d", C);
    GW_REQUIRE(p->mode == 0 || (coef != nullptr && net_b != nullptr), "gw_final_step: step mode needs coef and net_b");
    GW_REQUIRE(p->mode == 1 || eps_out != nullptr, "gw_final_step: forward mode needs eps_out");
    GW_REQUIRE(!(p->dc_weight > 0.0f) || p->y_dc != nullptr, "gw_final_step: dc_weight needs y_dc");
    GW_REQUIRE(dtype == GW_F32 || dtype == GW_BF16 || dtype == GW_DOTS, "gw_final_step: dtype %d", dtype);
    if (dtype == GW_DOTS)                                    // h = the head dots left by gw_conv_gn2 (stream_gn.cu)
        return final_step_dots(h, net_a, net_b, B, Cx, L, C, wf, bf, p, coef, step_ptr, noise, eps_out, x0_out, (cudaStream_t)stream);
    StepArgs a;
    a.mode = p->mode; a.cfg_both = p->cfg_both; a.selfcond = p->selfcond; a.pred_x0 = p->pred_x0;
    a.eps_scale = p->eps_scale; a.dc_weight = p->dc_weight; a.y_dc = p->y_dc; a.seed = p->seed; a.sample0 = p->sample0;
    a.rng = p->rng;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GW_BF16 && C == 64 && g_final_stream)       // HBM-streaming kernel (stream_gn.cu)
        return final_step_stream(h, net_a, net_b, B, Cx, L, wf, bf, p, coef, step_ptr, noise, eps_out, x0_out, st);
    dim3 grid(gw_cdiv(L, FS_TP), B);
    if (dtype == GW_F32)
        GW_CUDA(gw_launch_pdl(final_step_kernel<float>, grid, dim3(256), (size_t)(0), st, (const float*)h, net_a, net_b ? net_b : net_a, B, Cx, L, C, wf, bf, a,
                                                       coef, step_ptr, noise, eps_out, x0_out));
    else
        GW_CUDA(gw_launch_pdl(final_step_kernel<bf16>, grid, dim3(256), (size_t)(0), st, (const bf16*)h, net_a, net_b ? net_b : net_a, B, Cx, L, C, wf, bf, a,
                                                      coef, step_ptr, noise, eps_out, x0_out));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

__global__ void step_advance_kernel(int* p, int set_value) {
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x == 0 && blockIdx.x == 0) *p = set_value >= 0 ? set_value : *p + 1;
}
extern "C" int gw_step_advance(int* step_ptr, int set_value, void* stream) {
    GW_CUDA(gw_launch_pdl(step_advance_kernel, dim3(1), dim3(32), (size_t)(0), (cudaStream_t)stream, step_ptr, set_value));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------
// standard normals from the Philox stream (seed, global sample index, step): x_T of a chain is step 0
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) philox_normal_kernel(unsigned long long seed, long sample0, unsigned step, int L,
                                                            float* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.y;
    const int l4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (l4 >= L) return;
    float z[4];
    Philox::normal4(seed, (uint32_t)(sample0 + b), step, (uint32_t)(l4 >> 2), z);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (l4 + i < L) out[(size_t)b * L + l4 + i] = z[i];
}
extern "C" int gw_philox_normal(unsigned long long seed, long sample0, unsigned step, int B, int L, float* out, void* stream) {
    GW_REQUIRE(B > 0 && L > 0 && out != nullptr, "gw_philox_normal: sizes");
    GW_CUDA(gw_launch_pdl(philox_normal_kernel, dim3(gw_cdiv(gw_cdiv(L, 4), 256), B), dim3(256), (size_t)(0), (cudaStream_t)stream, seed, sample0, step, L, out));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------
// q_sample
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) q_sample_kernel(const float* __restrict__ x0, const int64_t* __restrict__ t,
                                                       const float* __restrict__ sab, const float* __restrict__ s1mab,
                                                       float* __restrict__ eps, int philox, unsigned long long seed,
                                                       long sample0, unsigned step, float clampv, float* __restrict__ net,
                                                       int Cx, int L) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.y;
    const int l4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (l4 >= L) return;
    const long tt = t[b];
    const float a = sab[tt], m = s1mab[tt];
    float z[4];
    if (philox) {
        Philox::normal4(seed, (uint32_t)(sample0 + b), step, (uint32_t)(l4 >> 2), z);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int l = l4 + i;
        if (l >= L) break;
        float e;
        if (philox) {
            e = z[i];
            eps[(size_t)b * L + l] = e;
        } else {
            e = eps[(size_t)b * L + l];
        }
        float v = a * x0[(size_t)b * L + l] + m * e;
        if (clampv > 0.0f) v = fminf(fmaxf(v, -clampv), clampv);
        net[(size_t)b * Cx * L + l] = v;
    }
}

extern "C" int gw_q_sample(const float* x0, const int64_t* t, const float* sqrt_ab, const float* sqrt_1mab, float* eps,
                           int philox, unsigned long long seed, long sample0, unsigned step, float clampv, float* net,
                           int B, int Cx, int L, void* stream) {
    GW_REQUIRE(B > 0 && L > 0 && Cx > 0, "gw_q_sample: sizes");
    dim3 grid(gw_cdiv(gw_cdiv(L, 4), 256), B);
    GW_CUDA(gw_launch_pdl(q_sample_kernel, grid, dim3(256), (size_t)(0), (cudaStream_t)stream, x0, t, sqrt_ab, sqrt_1mab, eps, philox, seed, sample0, step, clampv,
                                                            net, Cx, L));
    GW_LAUNCH_CHECK();
    return GW_OK;
}
