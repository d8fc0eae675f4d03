// Shape-generic CUDA-core kernels for the architectures the fast kernels do not cover (sm_100a):
//   UNet1D(base_ch = anything, kernel = 3 | 5 | 7, ...)   (models.py:78-88 allows them; the CLI defaults are base_ch = 64, kernel = 3)
// The tcgen05 / streaming kernels vectorise over 64-channel rows and hard-wire three taps; these kernels make no such
// assumption: any channel count (GroupNorm with gcd(8, C) groups, models.py:163), any odd kernel size, fp32 arithmetic on fp32
// or bf16 channels-last storage.  They are the exact path for non-default models, not a speed path: one thread per output
// element, reductions through shared memory and fp32 atomics on the (small) parameter gradients.
//   forward : gw_gen_conv -> gw_gen_gn_stats -> gw_gen_gn_apply ... -> gw_gen_final (+ CFG / DDIM update, step_update.cuh)
//   backward: gw_gen_final_bwd, gw_gen_gn_bwd, gw_gen_wgrad, gw_gen_weight_dgrad + gw_gen_conv (dgrad), gw_split_cat_grad (backward.cu)
#include "common.cuh"
#include "../../include/gwb200.h"
#include "step_update.cuh"

#define GEN_MAX_CC 8
#define GEN_MAX_K 7

template <typename T>
__device__ __forceinline__ float gen_ld(const T* p) { return to_f(*p); }
__device__ __forceinline__ void gen_st(float* p, float v) { *p = v; }
__device__ __forceinline__ void gen_st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// ------------------------------------------------------------------------------------------------ conv (any Cin, Cout, K)
// in(b, l, ci): channels-last src0 [B, L0, C0] (optionally nearest-upsampled: row l >> 1, zero beyond L0 -- models.py:217-220)
// followed by src1 [B, L, C1]; or (src0 == NULL) the channel-first fp32 network input x [B, Cx, L].  Zero padding K/2.
template <typename T>
__global__ void __launch_bounds__(256) gen_conv_kernel(const T* __restrict__ src0, int C0, int L0, int up0, const T* __restrict__ src1,
                                                       int C1, const float* __restrict__ xa, const float* __restrict__ xb,
                                                       const int* __restrict__ step_ptr, int Cx, int B, int L,
                                                       const float* __restrict__ w, const float* __restrict__ bias, int Cout, int K,
                                                       T* __restrict__ out) {
    const long n = (long)B * L * Cout;
    const int Cin = src0 != nullptr ? C0 + C1 : Cx;
    const int pad = K / 2;
    const float* x = nullptr;
    if (src0 == nullptr) {
        const int step = step_ptr != nullptr ? *step_ptr : 0;
        x = (step & 1) ? xb : xa;
    }
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        const long bl = i / Cout;
        const int l = (int)(bl % L), b = (int)(bl / L);
        float acc = bias != nullptr ? bias[co] : 0.0f;
        const float* wr = w + (size_t)co * Cin * K;
        for (int k = 0; k < K; ++k) {
            const int ll = l + k - pad;
            if (ll < 0 || ll >= L) continue;
            if (x != nullptr) {
                for (int ci = 0; ci < Cin; ++ci) acc = fmaf(x[((size_t)b * Cx + ci) * L + ll], wr[ci * K + k], acc);
            } else {
                const int ls = up0 ? (ll >> 1) : ll;
                if (ls < L0) {
                    const T* r0 = src0 + ((size_t)b * L0 + ls) * C0;
                    for (int ci = 0; ci < C0; ++ci) acc = fmaf(gen_ld(r0 + ci), wr[ci * K + k], acc);
                }
                if (C1 > 0) {
                    const T* r1 = src1 + ((size_t)b * L + ll) * C1;
                    for (int ci = 0; ci < C1; ++ci) acc = fmaf(gen_ld(r1 + ci), wr[(C0 + ci) * K + k], acc);
                }
            }
        }
        gen_st(out + i, acc);
    }
}

extern "C" int gw_gen_conv(const void* src0, int C0, int L0, int up0, const void* src1, int C1, const float* x, const float* x_alt,
                           const int* step_ptr, int Cx, int B, int L, const float* w, const float* bias, int Cout, int K, void* out,
                           int dtype, void* stream) {
    GW_REQUIRE((src0 != nullptr) != (x != nullptr), "gw_gen_conv: exactly one of src0 / x");
    GW_REQUIRE(K >= 1 && K <= GEN_MAX_K && (K & 1) == 1, "gw_gen_conv: kernel size %d (odd, <= %d)", K, GEN_MAX_K);
    GW_REQUIRE(B > 0 && L > 0 && Cout > 0 && w != nullptr && out != nullptr, "gw_gen_conv: sizes");
    GW_REQUIRE((src1 != nullptr) == (C1 > 0) && (src0 == nullptr || C0 > 0), "gw_gen_conv: channels");
    GW_REQUIRE(dtype == GW_F32 || dtype == GW_BF16, "gw_gen_conv: dtype %d", dtype);
    const long n = (long)B * L * Cout;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 32) grid = 148 * 32;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GW_F32)
        gen_conv_kernel<float><<<grid, 256, 0, st>>>((const float*)src0, C0, L0, up0, (const float*)src1, C1, x, x_alt ? x_alt : x, step_ptr,
                                                      Cx, B, L, w, bias, Cout, K, (float*)out);
    else
        gen_conv_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)src0, C0, L0, up0, (const bf16*)src1, C1, x, x_alt ? x_alt : x, step_ptr,
                                                     Cx, B, L, w, bias, Cout, K, (bf16*)out);
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// w'[ci][co][k] = w[co][ci][K-1-k]: the conv whose forward is the input gradient of the original ('same' padding, stride 1)
__global__ void gen_weight_dgrad_kernel(const float* __restrict__ w, int Cout, int Cin, int K, float* __restrict__ wt) {
    const int n = Cout * Cin * K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int k = i % K, co = (i / K) % Cout, ci = i / (K * Cout);
        wt[i] = w[((size_t)co * Cin + ci) * K + (K - 1 - k)];
    }
}
extern "C" int gw_gen_weight_dgrad(const float* w, int Cout, int Cin, int K, float* wt, void* stream) {
    GW_REQUIRE(w && wt && Cout > 0 && Cin > 0 && K >= 1 && K <= GEN_MAX_K, "gw_gen_weight_dgrad: arguments");
    gen_weight_dgrad_kernel<<<gw_cdiv(Cout * Cin * K, 256), 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, K, wt);
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------ GroupNorm statistics
__device__ __forceinline__ double gen_blk_sum(double v, double* red) {
    v = warp_sum_d(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    return t;
}
// one CTA per (sample, group): mean and rstd (biased variance, eps = 1e-5) of the stored values
template <typename T>
__global__ void __launch_bounds__(256) gen_gn_stats_kernel(const T* __restrict__ raw, int L, int C, int groups, float* __restrict__ stats) {
    __shared__ double red[8];
    const int b = blockIdx.x / groups, g = blockIdx.x % groups, cg = C / groups;
    const long n = (long)L * cg;
    double s1 = 0.0, s2 = 0.0;
    for (long i = threadIdx.x; i < n; i += 256) {
        const int l = (int)(i / cg), c = g * cg + (int)(i % cg);
        const double v = (double)gen_ld(raw + ((size_t)b * L + l) * C + c);
        s1 += v;
        s2 += v * v;
    }
    s1 = gen_blk_sum(s1, red);
    s2 = gen_blk_sum(s2, red);
    if (threadIdx.x == 0) {
        const double mean = s1 / (double)n;
        double var = s2 / (double)n - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[((size_t)b * 8 + g) * 2 + 0] = (float)mean;
        stats[((size_t)b * 8 + g) * 2 + 1] = (float)(1.0 / sqrt(var + 1e-5));
    }
}
extern "C" int gw_gen_gn_stats(const void* raw, int B, int L, int C, int groups, int dtype, float* stats, void* stream) {
    GW_REQUIRE(raw && stats && B > 0 && L > 0 && groups >= 1 && groups <= 8 && C % groups == 0, "gw_gen_gn_stats: C=%d groups=%d", C, groups);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GW_F32) gen_gn_stats_kernel<float><<<B * groups, 256, 0, st>>>((const float*)raw, L, C, groups, stats);
    else gen_gn_stats_kernel<bf16><<<B * groups, 256, 0, st>>>((const bf16*)raw, L, C, groups, stats);
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------ GroupNorm apply (forward)
// out = (silu(GN(raw)) + cond 1x1 conv) * (1 + gamma_t) + beta_t; pooled = avg_pool1d(out, 2, 2)  (models.py:160-173, 188-193, 205-208)
struct GenGnArgs {
    const float* stats; const float* gn_w; const float* gn_b; const float* cond; const float* wc; const float* bc; const float* film;
    const int* step_ptr;
    long film_b_stride, film_step_stride;
    int film_off, B, L, C, groups, Cc;
};
template <typename T>
__global__ void __launch_bounds__(256) gen_gn_apply_kernel(const T* __restrict__ raw, GenGnArgs A, T* __restrict__ out, T* __restrict__ pooled) {
    const int Lp = (A.L + 1) / 2;                          // row pairs (the last one may be a single row)
    const long n = (long)A.B * Lp * A.C;
    const int cg = A.C / A.groups;
    const int step = A.step_ptr != nullptr ? *A.step_ptr : 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % A.C);
        const long bp = i / A.C;
        const int p = (int)(bp % Lp), b = (int)(bp / Lp);
        const float mean = A.stats[((size_t)b * 8 + c / cg) * 2], rstd = A.stats[((size_t)b * 8 + c / cg) * 2 + 1];
        const float* fr = A.film + (size_t)step * A.film_step_stride + (size_t)b * A.film_b_stride + A.film_off;
        const float g1 = 1.0f + fr[c], be = fr[A.C + c];
        const float gw = A.gn_w[c], gb = A.gn_b[c];
        float o[2] = {0.0f, 0.0f};
        for (int h = 0; h < 2; ++h) {
            const int l = 2 * p + h;
            if (l >= A.L) break;
            const float x = gen_ld(raw + ((size_t)b * A.L + l) * A.C + c);
            const float y = (x - mean) * rstd * gw + gb;
            float u = silu_f<false>(y);
            if (A.Cc > 0) {
                float cb = A.bc[c];
                const float* cr = A.cond + ((size_t)b * A.L + l) * A.Cc;
                for (int j = 0; j < A.Cc; ++j) cb = fmaf(A.wc[c * A.Cc + j], cr[j], cb);
                u += cb;
            }
            o[h] = fmaf(u, g1, be);
            gen_st(out + ((size_t)b * A.L + l) * A.C + c, o[h]);
        }
        if (pooled != nullptr && 2 * p + 1 < A.L) gen_st(pooled + ((size_t)b * (A.L / 2) + p) * A.C + c, 0.5f * (o[0] + o[1]));
    }
}
extern "C" int gw_gen_gn_apply(const void* raw, const float* stats, int B, int L, int C, int groups, const float* gn_w, const float* gn_b,
                               const float* cond, int Cc, const float* wc, const float* bc, const float* film, int film_off,
                               long film_b_stride, long film_step_stride, const int* step_ptr, void* out, void* pooled, int dtype,
                               void* stream) {
    GW_REQUIRE(raw && stats && gn_w && gn_b && film && out && B > 0 && L > 0 && groups >= 1 && groups <= 8 && C % groups == 0,
               "gw_gen_gn_apply: arguments (C=%d groups=%d)", C, groups);
    GW_REQUIRE(Cc >= 0 && Cc <= GEN_MAX_CC && (Cc == 0 || (cond && wc && bc)), "gw_gen_gn_apply: Cc=%d", Cc);
    GenGnArgs A{stats, gn_w, gn_b, cond, wc, bc, film, step_ptr, film_b_stride, film_step_stride, film_off, B, L, C, groups, Cc};
    const long n = (long)B * ((L + 1) / 2) * C;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 32) grid = 148 * 32;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GW_F32) gen_gn_apply_kernel<float><<<grid, 256, 0, st>>>((const float*)raw, A, (float*)out, (float*)pooled);
    else gen_gn_apply_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)raw, A, (bf16*)out, (bf16*)pooled);
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------ head conv (+ reverse step)
// eps[b, l] = bf + sum_{c < C} sum_k wf[c, k] h[b, l + k - K/2, c] + sum_k wf[C, k] x_t[b, l + k - K/2]   (models.py:227-230),
// mode 1: CFG combine + DDIM / DDPM update of the position (step_update.cuh), as gw_final_step.
template <typename T>
__global__ void __launch_bounds__(256) gen_final_kernel(const T* __restrict__ h, const float* __restrict__ net_a,
                                                        const float* __restrict__ net_b, int B, int Cx, int L, int C, int K,
                                                        const float* __restrict__ wf, const float* __restrict__ bf, FssArgs p,
                                                        const float* __restrict__ coef, const int* __restrict__ step_ptr,
                                                        const float* __restrict__ noise, float* __restrict__ eps_out,
                                                        float* __restrict__ x0_out) {
    const int b = blockIdx.y, l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const int step = step_ptr != nullptr ? *step_ptr : 0;
    const float* net_in = (step & 1) ? net_b : net_a;
    float* net_out = const_cast<float*>((step & 1) ? net_a : net_b);
    const int n_half = (p.mode == 1 && p.cfg_both) ? 2 : 1;
    const int pad = K / 2;
    float outv[2] = {0.0f, 0.0f};
    for (int hf = 0; hf < n_half; ++hf) {
        const int bb = b + hf * B;
        float acc = bf[0];
        for (int k = 0; k < K; ++k) {
            const int ll = l + k - pad;
            if (ll < 0 || ll >= L) continue;
            const T* hr = h + ((size_t)bb * L + ll) * C;
            for (int c = 0; c < C; ++c) acc = fmaf(gen_ld(hr + c), wf[c * K + k], acc);
            acc = fmaf(net_in[(size_t)bb * Cx * L + ll], wf[C * K + k], acc);
        }
        outv[hf] = acc;
    }
    if (p.mode == 0) {
        eps_out[(size_t)b * L + l] = outv[0];
        return;
    }
    const float* cfp = coef + (size_t)step * 16;
    FssCoef cf;
    cf.c_s1mab = cfp[0]; cf.c_sab = cfp[1]; cf.c_sabp = cfp[2]; cf.c_dir = cfp[3]; cf.c_sig = cfp[4]; cf.c_w = cfp[5];
    cf.use = (int)cfp[6]; cf.last = (int)cfp[7]; cf.draw = (int)cfp[8]; cf.c_s1mab_cl = cfp[9];
    float zu = 0.0f;
    if (noise == nullptr && !cf.last && cf.c_sig > 0.0f) {
        float z4[4];
        const unsigned long long sd = p.rng != nullptr ? p.rng[0] : p.seed;
        const long s0 = p.rng != nullptr ? (long)p.rng[1] : p.sample0;
        Philox::normal4(sd, (uint32_t)(s0 + b), (uint32_t)step + 1u, (uint32_t)(l >> 2), z4);
        zu = z4[l & 3];
    }
    fss_update(p, cf, outv[0], outv[1], net_in[(size_t)b * Cx * L + l], zu, noise, net_out, n_half, b, B, Cx, L, l, eps_out, x0_out);
}
extern "C" int gw_gen_final(const void* h, int dtype, const float* net_a, const float* net_b, int B, int Cx, int L, int C, int K,
                            const float* wf, const float* bf, const gw_step_params* p, const float* coef, const int* step_ptr,
                            const float* noise, float* eps_out, float* x0_out, void* stream) {
    GW_REQUIRE(h && net_a && wf && bf && p && B > 0 && L > 0 && C > 0 && K >= 1 && K <= GEN_MAX_K && (K & 1), "gw_gen_final: arguments");
    GW_REQUIRE(p->mode == 0 || (coef != nullptr && net_b != nullptr), "gw_gen_final: step mode needs coef and net_b");
    GW_REQUIRE(p->mode == 1 || eps_out != nullptr, "gw_gen_final: forward mode needs eps_out");
    GW_REQUIRE(!(p->dc_weight > 0.0f) || p->y_dc != nullptr, "gw_gen_final: dc_weight needs y_dc");
    FssArgs a;
    a.mode = p->mode; a.cfg_both = p->cfg_both; a.selfcond = p->selfcond; a.pred_x0 = p->pred_x0;
    a.eps_scale = p->eps_scale; a.dc_weight = p->dc_weight; a.y_dc = p->y_dc; a.seed = p->seed; a.sample0 = p->sample0; a.rng = p->rng;
    a.advance = nullptr;
    dim3 grid(gw_cdiv(L, 256), B);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GW_F32)
        gen_final_kernel<float><<<grid, 256, 0, st>>>((const float*)h, net_a, net_b ? net_b : net_a, B, Cx, L, C, K, wf, bf, a, coef, step_ptr,
                                                       noise, eps_out, x0_out);
    else
        gen_final_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)h, net_a, net_b ? net_b : net_a, B, Cx, L, C, K, wf, bf, a, coef, step_ptr,
                                                      noise, eps_out, x0_out);
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ================================================================================================ backward
// head conv backward: d_h[b, l, c] = sum_k wf[c, k] d_eps[b, l - k + K/2];  dWf[c, k] += sum_{b, l} d_eps[b, l] hcat[b, l + k - K/2, c]
// (hcat = [h | x_t]);  dbf += sum d_eps.
template <typename T>
__global__ void __launch_bounds__(256) gen_final_bwd_dh_kernel(const float* __restrict__ d_eps, int B, int L, int C, int K,
                                                               const float* __restrict__ wf, T* __restrict__ d_h) {
    const long n = (long)B * L * C;
    const int pad = K / 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long bl = i / C;
        const int l = (int)(bl % L), b = (int)(bl / L);
        float acc = 0.0f;
        for (int k = 0; k < K; ++k) {
            const int lo = l - k + pad;
            if (lo >= 0 && lo < L) acc = fmaf(wf[c * K + k], d_eps[(size_t)b * L + lo], acc);
        }
        gen_st(d_h + i, acc);
    }
}
// one CTA per (channel c of hcat, tap k); c == C is the x_t channel; the CTA (0, 0) also reduces dbf
template <typename T>
__global__ void __launch_bounds__(256) gen_final_bwd_w_kernel(const float* __restrict__ d_eps, const T* __restrict__ h,
                                                              const float* __restrict__ net, int B, int Cx, int L, int C, int K,
                                                              float* __restrict__ dWf, float* __restrict__ dbf) {
    __shared__ double red[8];
    const int c = blockIdx.x, k = blockIdx.y, pad = K / 2;
    double s = 0.0, sb = 0.0;
    for (long i = threadIdx.x; i < (long)B * L; i += 256) {
        const int b = (int)(i / L), l = (int)(i % L);
        const float de = d_eps[i];
        sb += (double)de;
        const int ll = l + k - pad;
        if (ll < 0 || ll >= L) continue;
        const float v = c < C ? gen_ld(h + ((size_t)b * L + ll) * C + c) : net[(size_t)b * Cx * L + ll];
        s += (double)de * (double)v;
    }
    s = gen_blk_sum(s, red);
    if (threadIdx.x == 0) dWf[c * K + k] += (float)s;
    if (c == 0 && k == 0) {
        sb = gen_blk_sum(sb, red);
        if (threadIdx.x == 0) dbf[0] += (float)sb;
    }
}
extern "C" int gw_gen_final_bwd(const float* d_eps, const void* h, int dtype, const float* net, int B, int Cx, int L, int C, int K,
                                const float* wf, void* d_h, float* dWf, float* dbf, void* stream) {
    GW_REQUIRE(d_eps && h && net && wf && d_h && dWf && dbf && B > 0 && L > 0 && C > 0 && K >= 1 && K <= GEN_MAX_K, "gw_gen_final_bwd: arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const long n = (long)B * L * C;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 32) grid = 148 * 32;
    if (dtype == GW_F32) {
        gen_final_bwd_dh_kernel<float><<<grid, 256, 0, st>>>(d_eps, B, L, C, K, wf, (float*)d_h);
        gen_final_bwd_w_kernel<float><<<dim3(C + 1, K), 256, 0, st>>>(d_eps, (const float*)h, net, B, Cx, L, C, K, dWf, dbf);
    } else {
        gen_final_bwd_dh_kernel<bf16><<<grid, 256, 0, st>>>(d_eps, B, L, C, K, wf, (bf16*)d_h);
        gen_final_bwd_w_kernel<bf16><<<dim3(C + 1, K), 256, 0, st>>>(d_eps, (const bf16*)h, net, B, Cx, L, C, K, dWf, dbf);
    }
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------ GroupNorm block backward
// Forward: y = xhat gn_w + gn_b (xhat = (raw - mean) rstd), a = silu(y), u = a + cb, out = u (1 + gamma) + beta.
// Given d_out (and d_pool: the gradient wrt the pooled output, spread 1/2 : 1/2 over its row pair):
//   pass 1, one CTA per (b, c): S[0] = sum d_o, S[1] = sum d_o u, S[2] = sum d_y, S[3] = sum d_y xhat, S[4] = sum xhat,
//           S[5 + j] = sum d_u cond_j                      (d_u = d_o (1 + gamma), d_y = d_u silu'(y))
//   pass 2, per element: d_raw = rstd gn_w (d_y - m1 - xhat m2),  m1 = mean_group(d_y gn_w), m2 = mean_group(d_y gn_w xhat)
//   pass 3, per channel: parameter gradients (sums over b), FiLM gradient rows, conv-bias gradient (analytic from the sums)
#define GEN_NS (5 + GEN_MAX_CC)
struct GenBwdArgs {
    const float* stats; const float* gn_w; const float* gn_b; const float* cond; const float* wc; const float* bc; const float* film;
    long film_b_stride;
    int film_off, B, L, C, groups, Cc;
};
template <typename T>
__device__ __forceinline__ float gen_dout(const T* d_out, const T* d_pool, int b, int l, int c, int L, int C) {
    float d = d_out != nullptr ? gen_ld(d_out + ((size_t)b * L + l) * C + c) : 0.0f;
    if (d_pool != nullptr && l < 2 * (L / 2)) d += 0.5f * gen_ld(d_pool + ((size_t)b * (L / 2) + (l >> 1)) * C + c);
    return d;
}
__device__ __forceinline__ float gen_dsilu(float y) {
    const float s = 1.0f / (1.0f + expf(-y));
    return s * (1.0f + y * (1.0f - s));
}
template <typename T>
__global__ void __launch_bounds__(256) gen_gn_bwd_sums_kernel(const T* __restrict__ raw, const T* __restrict__ d_out,
                                                              const T* __restrict__ d_pool, GenBwdArgs A, float* __restrict__ S) {
    __shared__ double red[8];
    const int b = blockIdx.x / A.C, c = blockIdx.x % A.C, cg = A.C / A.groups;
    const float mean = A.stats[((size_t)b * 8 + c / cg) * 2], rstd = A.stats[((size_t)b * 8 + c / cg) * 2 + 1];
    const float* fr = A.film + (size_t)b * A.film_b_stride + A.film_off;
    const float g1 = 1.0f + fr[c], gw = A.gn_w[c], gb = A.gn_b[c];
    double s[GEN_NS];
    for (int j = 0; j < GEN_NS; ++j) s[j] = 0.0;
    for (int l = threadIdx.x; l < A.L; l += 256) {
        const float x = gen_ld(raw + ((size_t)b * A.L + l) * A.C + c);
        const float xh = (x - mean) * rstd;
        const float y = xh * gw + gb;
        float u = silu_f<false>(y);
        const float* cr = A.Cc > 0 ? A.cond + ((size_t)b * A.L + l) * A.Cc : nullptr;
        if (A.Cc > 0) {
            float cb = A.bc[c];
            for (int j = 0; j < A.Cc; ++j) cb = fmaf(A.wc[c * A.Cc + j], cr[j], cb);
            u += cb;
        }
        const float d_o = gen_dout(d_out, d_pool, b, l, c, A.L, A.C);
        const float d_u = d_o * g1;
        const float d_y = d_u * gen_dsilu(y);
        s[0] += d_o; s[1] += (double)d_o * u; s[2] += d_y; s[3] += (double)d_y * xh; s[4] += xh;
        for (int j = 0; j < A.Cc; ++j) s[5 + j] += (double)d_u * cr[j];
    }
    // sum of d_u itself (d_bc) = S[0] (1 + gamma): no extra slot needed
    for (int j = 0; j < 5 + A.Cc; ++j) {
        const double t = gen_blk_sum(s[j], red);
        if (threadIdx.x == 0) S[((size_t)b * A.C + c) * GEN_NS + j] = (float)t;
    }
}
template <typename T>
__global__ void __launch_bounds__(256) gen_gn_bwd_apply_kernel(const T* __restrict__ raw, const T* __restrict__ d_out,
                                                               const T* __restrict__ d_pool, GenBwdArgs A, const float* __restrict__ S,
                                                               T* __restrict__ d_raw) {
    const long n = (long)A.B * A.L * A.C;
    const int cg = A.C / A.groups;
    const float inv_n = 1.0f / ((float)cg * (float)A.L);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % A.C);
        const long bl = i / A.C;
        const int l = (int)(bl % A.L), b = (int)(bl / A.L);
        const int g = c / cg;
        const float mean = A.stats[((size_t)b * 8 + g) * 2], rstd = A.stats[((size_t)b * 8 + g) * 2 + 1];
        float m1 = 0.0f, m2 = 0.0f;                       // group means of d_xhat and d_xhat * xhat (d_xhat = d_y gn_w)
        for (int cc = g * cg; cc < (g + 1) * cg; ++cc) {
            const float* sp = S + ((size_t)b * A.C + cc) * GEN_NS;
            m1 = fmaf(sp[2], A.gn_w[cc], m1);
            m2 = fmaf(sp[3], A.gn_w[cc], m2);
        }
        m1 *= inv_n; m2 *= inv_n;
        const float* fr = A.film + (size_t)b * A.film_b_stride + A.film_off;
        const float gw = A.gn_w[c];
        const float x = gen_ld(raw + i);
        const float xh = (x - mean) * rstd;
        const float y = xh * gw + A.gn_b[c];
        const float d_o = gen_dout(d_out, d_pool, b, l, c, A.L, A.C);
        const float d_xh = d_o * (1.0f + fr[c]) * gen_dsilu(y) * gw;
        gen_st(d_raw + i, rstd * (d_xh - m1 - xh * m2));
    }
}
// one thread per channel: reduce the per-(b, c) sums over b into the parameter gradients (+=) and write the FiLM gradient rows
__global__ void gen_gn_bwd_params_kernel(GenBwdArgs A, const float* __restrict__ S, float* __restrict__ dfilm, long dfilm_stride,
                                         float* __restrict__ d_gn_w, float* __restrict__ d_gn_b, float* __restrict__ d_wc,
                                         float* __restrict__ d_bc, float* __restrict__ d_bias) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= A.C) return;
    const int cg = A.C / A.groups, g = c / cg;
    const float inv_n = 1.0f / ((float)cg * (float)A.L);
    double a_gw = 0.0, a_gb = 0.0, a_bc = 0.0, a_bias = 0.0, a_wc[GEN_MAX_CC];
    for (int j = 0; j < GEN_MAX_CC; ++j) a_wc[j] = 0.0;
    for (int b = 0; b < A.B; ++b) {
        const float* sp = S + ((size_t)b * A.C + c) * GEN_NS;
        const float* fr = A.film + (size_t)b * A.film_b_stride + A.film_off;
        dfilm[(size_t)b * dfilm_stride + A.film_off + c] = sp[1];                 // d gamma_t = sum d_o u
        dfilm[(size_t)b * dfilm_stride + A.film_off + A.C + c] = sp[0];           // d beta_t  = sum d_o
        a_gw += sp[3]; a_gb += sp[2];
        a_bc += (double)sp[0] * (1.0 + (double)fr[c]);
        for (int j = 0; j < A.Cc; ++j) a_wc[j] += sp[5 + j];
        // sum_l d_raw = rstd (gn_w S2 - L m1 - S4 m2) with the group means of this sample
        float m1 = 0.0f, m2 = 0.0f;
        for (int cc = g * cg; cc < (g + 1) * cg; ++cc) {
            const float* sq = S + ((size_t)b * A.C + cc) * GEN_NS;
            m1 = fmaf(sq[2], A.gn_w[cc], m1);
            m2 = fmaf(sq[3], A.gn_w[cc], m2);
        }
        m1 *= inv_n; m2 *= inv_n;
        const float rstd = A.stats[((size_t)b * 8 + g) * 2 + 1];
        a_bias += (double)rstd * ((double)A.gn_w[c] * sp[2] - (double)A.L * m1 - (double)sp[4] * m2);
    }
    d_gn_w[c] += (float)a_gw;
    d_gn_b[c] += (float)a_gb;
    if (A.Cc > 0) {
        d_bc[c] += (float)a_bc;
        for (int j = 0; j < A.Cc; ++j) d_wc[c * A.Cc + j] += (float)a_wc[j];
    }
    d_bias[c] += (float)a_bias;
}
extern "C" long gw_gen_gn_bwd_scratch_floats(int B, int C) { return (long)B * C * GEN_NS; }
extern "C" int gw_gen_gn_bwd(const void* raw, const float* stats, int B, int L, int C, int groups, const float* gn_w, const float* gn_b,
                             const float* cond, int Cc, const float* wc, const float* bc, const float* film, int film_off,
                             long film_b_stride, const void* d_out, const void* d_pool, int dtype, float* scratch, float* dfilm,
                             long dfilm_stride, void* d_raw, float* d_gn_w, float* d_gn_b, float* d_wc, float* d_bc, float* d_bias,
                             void* stream) {
    GW_REQUIRE(raw && stats && gn_w && gn_b && film && scratch && dfilm && d_raw && d_gn_w && d_gn_b && d_bias && (d_out || d_pool),
               "gw_gen_gn_bwd: null pointer");
    GW_REQUIRE(B > 0 && L > 0 && groups >= 1 && groups <= 8 && C % groups == 0 && Cc >= 0 && Cc <= GEN_MAX_CC, "gw_gen_gn_bwd: sizes");
    GW_REQUIRE(Cc == 0 || (cond && wc && bc && d_wc && d_bc), "gw_gen_gn_bwd: cond pointers");
    GenBwdArgs A{stats, gn_w, gn_b, cond, wc, bc, film, film_b_stride, film_off, B, L, C, groups, Cc};
    cudaStream_t st = (cudaStream_t)stream;
    const long n = (long)B * L * C;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 32) grid = 148 * 32;
    if (dtype == GW_F32) {
        gen_gn_bwd_sums_kernel<float><<<B * C, 256, 0, st>>>((const float*)raw, (const float*)d_out, (const float*)d_pool, A, scratch);
        gen_gn_bwd_apply_kernel<float><<<grid, 256, 0, st>>>((const float*)raw, (const float*)d_out, (const float*)d_pool, A, scratch, (float*)d_raw);
    } else {
        gen_gn_bwd_sums_kernel<bf16><<<B * C, 256, 0, st>>>((const bf16*)raw, (const bf16*)d_out, (const bf16*)d_pool, A, scratch);
        gen_gn_bwd_apply_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)raw, (const bf16*)d_out, (const bf16*)d_pool, A, scratch, (bf16*)d_raw);
    }
    GW_LAUNCH_CHECK();
    gen_gn_bwd_params_kernel<<<gw_cdiv(C, 128), 128, 0, st>>>(A, scratch, dfilm, dfilm_stride, d_gn_w, d_gn_b, d_wc, d_bc, d_bias);
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[co, ci, k] += sum_{b, l} d_raw[b, l, co] in(b, l + k - K/2, ci); one CTA per (co, ci), K accumulators per thread
template <typename T>
__global__ void __launch_bounds__(256) gen_wgrad_kernel(const T* __restrict__ src0, int C0, int L0, int up0, const T* __restrict__ src1,
                                                        int C1, const float* __restrict__ x, int Cx, const T* __restrict__ d_raw, int B,
                                                        int L, int Cout, int K, float* __restrict__ dW) {
    __shared__ double red[8];
    const int co = blockIdx.x, ci = blockIdx.y, pad = K / 2;
    const int Cin = src0 != nullptr ? C0 + C1 : Cx;
    double s[GEN_MAX_K];
    for (int k = 0; k < GEN_MAX_K; ++k) s[k] = 0.0;
    for (long i = threadIdx.x; i < (long)B * L; i += 256) {
        const int b = (int)(i / L), l = (int)(i % L);
        float v = 0.0f;                                   // in(b, l, ci)
        if (x != nullptr) v = x[((size_t)b * Cx + ci) * L + l];
        else if (ci < C0) {
            const int ls = up0 ? (l >> 1) : l;
            if (ls < L0) v = gen_ld(src0 + ((size_t)b * L0 + ls) * C0 + ci);
        } else v = gen_ld(src1 + ((size_t)b * L + l) * C1 + (ci - C0));
        if (v == 0.0f) continue;
        for (int k = 0; k < K; ++k) {                     // in(l) meets d_raw(l - k + pad) under tap k
            const int lo = l - k + pad;
            if (lo >= 0 && lo < L) s[k] += (double)v * (double)gen_ld(d_raw + ((size_t)b * L + lo) * Cout + co);
        }
    }
    for (int k = 0; k < K; ++k) {
        const double t = gen_blk_sum(s[k], red);
        if (threadIdx.x == 0) dW[((size_t)co * Cin + ci) * K + k] += (float)t;
    }
}
extern "C" int gw_gen_wgrad(const void* src0, int C0, int L0, int up0, const void* src1, int C1, const float* x, int Cx,
                            const void* d_raw, int B, int L, int Cout, int K, int dtype, float* dW, void* stream) {
    GW_REQUIRE((src0 != nullptr) != (x != nullptr) && d_raw && dW && B > 0 && L > 0 && Cout > 0 && K >= 1 && K <= GEN_MAX_K,
               "gw_gen_wgrad: arguments");
    const int Cin = src0 != nullptr ? C0 + C1 : Cx;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GW_F32)
        gen_wgrad_kernel<float><<<dim3(Cout, Cin), 256, 0, st>>>((const float*)src0, C0, L0, up0, (const float*)src1, C1, x, Cx,
                                                                  (const float*)d_raw, B, L, Cout, K, dW);
    else
        gen_wgrad_kernel<bf16><<<dim3(Cout, Cin), 256, 0, st>>>((const bf16*)src0, C0, L0, up0, (const bf16*)src1, C1, x, Cx,
                                                                 (const bf16*)d_raw, B, L, Cout, K, dW);
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// d_cat [B, L, C0 + C1] (gradient wrt cat[upsample(h), skip]) -> d_h [B, L0, C0] (sum of the row pair; rows beyond L ignored),
// d_skip [B, L, C1]; any channel counts (backward.cu's gw_split_cat_grad needs multiples of 4)
template <typename T>
__global__ void __launch_bounds__(256) gen_split_cat_kernel(const T* __restrict__ d_cat, int B, int L, int C0, int L0, int C1,
                                                            T* __restrict__ d_h, T* __restrict__ d_skip) {
    const int C = C0 + C1;
    const long n0 = (long)B * L0 * C0, n1 = (long)B * L * C1;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1; i += (long)gridDim.x * blockDim.x) {
        if (i < n0) {
            const int c = (int)(i % C0);
            const long bm = i / C0;
            const int m = (int)(bm % L0), b = (int)(bm / L0);
            float a = 0.0f;
            for (int h = 0; h < 2; ++h) {
                const int l = 2 * m + h;
                if (l < L) a += gen_ld(d_cat + ((size_t)b * L + l) * C + c);
            }
            gen_st(d_h + i, a);
        } else {
            const long j = i - n0;
            const int c = (int)(j % C1);
            const long bl = j / C1;
            gen_st(d_skip + j, gen_ld(d_cat + (size_t)bl * C + C0 + c));
        }
    }
}
extern "C" int gw_gen_split_cat(const void* d_cat, int B, int L, int C0, int L0, int C1, void* d_h, void* d_skip, int dtype, void* stream) {
    GW_REQUIRE(d_cat && d_h && d_skip && B > 0 && L > 0 && L0 > 0 && C0 > 0 && C1 > 0, "gw_gen_split_cat: arguments");
    const long n = (long)B * L0 * C0 + (long)B * L * C1;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 32) grid = 148 * 32;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GW_F32) gen_split_cat_kernel<float><<<grid, 256, 0, st>>>((const float*)d_cat, B, L, C0, L0, C1, (float*)d_h, (float*)d_skip);
    else gen_split_cat_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)d_cat, B, L, C0, L0, C1, (bf16*)d_h, (bf16*)d_skip);
    GW_LAUNCH_CHECK();
    return GW_OK;
}
