// Training-step kernels around the network: input packing (q_sample + CFG dropout + clamp, train.py:349-407),
// self-conditioning estimate (train.py:40-51), time-MLP / tproj backward (models.py:105-109, 137-142),
// and the fused clip_grad_norm_ + AdamW + EMA update (train.py:445-455, 73-81) over flat parameter buffers.
#include "common.cuh"
#include "../../include/gwb200.h"

// ------------------------------------------------------------------------------------------------
// per-sample draws: t ~ U{t_min..T-1} (train.py:376) and the CFG-dropout coin (train.py:386), Philox keyed on the
// GLOBAL sample index so a W-rank run draws what the 1-rank run draws.
// ------------------------------------------------------------------------------------------------
__global__ void train_draws_kernel(unsigned long long seed, const int* __restrict__ step_ptr, long sample0, int B, int t_min,
                                   int T, float p_uncond, int64_t* __restrict__ t_out, float* __restrict__ drop_out) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t step = (uint32_t)(step_ptr ? *step_ptr : 0);
    uint32_t r[4];
    Philox::gen(seed, 0xfffffff0u, step, (uint32_t)(sample0 + b), 0x3c6ef372u, r);
    const int span = T - t_min;
    t_out[b] = (int64_t)t_min + (int64_t)(((unsigned long long)r[0] * (unsigned long long)span) >> 32);
    const float u = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    drop_out[b] = u < p_uncond ? 1.0f : 0.0f;
}
extern "C" int gw_train_draws(unsigned long long seed, const int* step_ptr, long sample0, int B, int t_min, int T,
                              float p_uncond, int64_t* t_out, float* drop_out, void* stream) {
    GW_REQUIRE(B > 0 && T > t_min && t_min >= 0, "gw_train_draws: B=%d t_min=%d T=%d", B, t_min, T);
    GW_CUDA(gw_launch_pdl(train_draws_kernel, dim3(gw_cdiv(B, 128)), dim3(128), (size_t)(0), (cudaStream_t)stream, seed, step_ptr, sample0, B, t_min, T, p_uncond, t_out, drop_out));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------
// network-input packing: net[b] = [ clamp(q_sample(clamp(clean), t, eps)) | clamp(y)*(1-drop), meta... | 0 ]
// (train.py:350-352, 379-398, 404-407).  cond [B, Cc, L] (channel 0 = y); eps read (philox == 0) or generated.
// drop_all != 0 zeroes every conditioning channel for dropped samples (dropout_y_only = False or no metadata);
// clamp_y != 0 clamps the y channel (the reference only does so on the y-only dropout branch, train.py:387).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) train_pack_kernel(const float* __restrict__ clean, const float* __restrict__ cond, int Cc,
                                                         const int64_t* __restrict__ t, const float* __restrict__ drop,
                                                         const float* __restrict__ sab, const float* __restrict__ s1mab,
                                                         float* __restrict__ eps, int philox, unsigned long long seed,
                                                         long sample0, const int* __restrict__ step_ptr, float clampv,
                                                         int clamp_y, int drop_all, float* __restrict__ net, int Cx, int L) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.y;
    const int l4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (l4 >= L) return;
    const long tt = t[b];
    const float a = sab[tt], m = s1mab[tt];
    const float keep = drop ? 1.0f - drop[b] : 1.0f;
    float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (philox) Philox::normal4(seed, (uint32_t)(sample0 + b), (uint32_t)(step_ptr ? *step_ptr : 0), (uint32_t)(l4 >> 2), z);
    float* nb = net + (size_t)b * Cx * L;
    if ((L & 3) == 0) {
        // whole quads: 16-byte loads / stores (the scalar path below touches every sector four times)
        const size_t o = (size_t)b * L + l4;
        float4 e4;
        if (philox) {
            e4 = make_float4(z[0], z[1], z[2], z[3]);
            *reinterpret_cast<float4*>(eps + o) = e4;
        } else {
            e4 = *reinterpret_cast<const float4*>(eps + o);
        }
        const float4 c4 = *reinterpret_cast<const float4*>(clean + o);
        const float ev[4] = {e4.x, e4.y, e4.z, e4.w}, xv[4] = {c4.x, c4.y, c4.z, c4.w};
        float vv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float x0 = xv[i];
            if (clampv > 0.0f) x0 = fminf(fmaxf(x0, -clampv), clampv);
            float v = a * x0 + m * ev[i];
            if (clampv > 0.0f) v = fminf(fmaxf(v, -clampv), clampv);
            vv[i] = v;
        }
        *reinterpret_cast<float4*>(nb + l4) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        for (int c = 0; c < Cc; ++c) {
            const float4 q = *reinterpret_cast<const float4*>(cond + ((size_t)b * Cc + c) * L + l4);
            float cv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (c == 0) {
                    if (clampv > 0.0f && clamp_y) cv[i] = fminf(fmaxf(cv[i], -clampv), clampv);
                    cv[i] *= keep;
                } else if (drop_all) {
                    cv[i] *= keep;
                }
            }
            *reinterpret_cast<float4*>(nb + (size_t)(1 + c) * L + l4) = make_float4(cv[0], cv[1], cv[2], cv[3]);
        }
        for (int c = 1 + Cc; c < Cx; ++c) *reinterpret_cast<float4*>(nb + (size_t)c * L + l4) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        return;
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
        float x0 = clean[(size_t)b * L + l];
        if (clampv > 0.0f) x0 = fminf(fmaxf(x0, -clampv), clampv);
        float v = a * x0 + m * e;
        if (clampv > 0.0f) v = fminf(fmaxf(v, -clampv), clampv);
        nb[l] = v;
        for (int c = 0; c < Cc; ++c) {
            float cv = cond[((size_t)b * Cc + c) * L + l];
            if (c == 0) {
                if (clampv > 0.0f && clamp_y) cv = fminf(fmaxf(cv, -clampv), clampv);   // train.py:387 uses the clamped y_norm, :398 cond_stack
                cv *= keep;
            } else if (drop_all) {
                cv *= keep;
            }
            nb[(size_t)(1 + c) * L + l] = cv;
        }
        for (int c = 1 + Cc; c < Cx; ++c) nb[(size_t)c * L + l] = 0.0f;
    }
}
extern "C" int gw_train_pack(const float* clean, const float* cond, int Cc, const int64_t* t, const float* drop,
                             const float* sqrt_ab, const float* sqrt_1mab, float* eps, int philox, unsigned long long seed,
                             long sample0, const int* step_ptr, float clampv, int clamp_y, int drop_all, float* net, int B, int Cx,
                             int L, void* stream) {
    GW_REQUIRE(B > 0 && L > 0 && Cx >= 1 + Cc && Cc >= 0, "gw_train_pack: sizes B=%d L=%d Cx=%d Cc=%d", B, L, Cx, Cc);
    dim3 grid(gw_cdiv(gw_cdiv(L, 4), 256), B);
    GW_CUDA(gw_launch_pdl(train_pack_kernel, grid, dim3(256), (size_t)(0), (cudaStream_t)stream, clean, cond, Cc, t, drop, sqrt_ab, sqrt_1mab, eps, philox, seed, sample0,
                                                              step_ptr, clampv, clamp_y, drop_all, net, Cx, L));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// collated batch -> stepper inputs in one pass (train.py:336-347 + the --t_multi repeat_interleave, :355-360): division by the
// per-sample sigma, y and the metadata channels stacked, every sample repeated K times.  out row b*K + r <- in row b.
__global__ void __launch_bounds__(256) batch_prepare_kernel(const float* __restrict__ clean_raw, const float* __restrict__ noisy_raw,
                                                            const float* __restrict__ sigma, const float* __restrict__ mask,
                                                            const float* __restrict__ meta, int Cm, int L, int K,
                                                            float* __restrict__ clean_out, float* __restrict__ cond_out,
                                                            float* __restrict__ mask_out) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.y;
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const float sg = sigma[b];
    const float c = __fdiv_rn(clean_raw[(size_t)b * L + l], sg), y = __fdiv_rn(noisy_raw[(size_t)b * L + l], sg);
    const float m = mask != nullptr ? mask[(size_t)b * L + l] : 1.0f;
    const int Cc = 1 + Cm;
    for (int r = 0; r < K; ++r) {
        const size_t o = (size_t)b * K + r;
        clean_out[o * L + l] = c;
        mask_out[o * L + l] = m;
        cond_out[o * Cc * L + l] = y;
        for (int j = 0; j < Cm; ++j) cond_out[(o * Cc + 1 + j) * L + l] = meta[((size_t)b * Cm + j) * L + l];
    }
}
extern "C" int gw_batch_prepare(const float* clean_raw, const float* noisy_raw, const float* sigma, const float* mask,
                                const float* meta, int Cm, int B0, int L, int K, float* clean_out, float* cond_out, float* mask_out,
                                void* stream) {
    GW_REQUIRE(clean_raw && noisy_raw && sigma && clean_out && cond_out && mask_out, "gw_batch_prepare: null pointer");
    GW_REQUIRE(B0 > 0 && L > 0 && K >= 1 && Cm >= 0 && (Cm == 0 || meta != nullptr), "gw_batch_prepare: sizes");
    GW_CUDA(gw_launch_pdl(batch_prepare_kernel, dim3(gw_cdiv(L, 256), B0), dim3(256), (size_t)(0), (cudaStream_t)stream, clean_raw, noisy_raw, sigma, mask, meta, Cm, L, K,
                                                                                  clean_out, cond_out, mask_out));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// self-conditioning estimate x0_hat = (x_t - sqrt(1-ab_t) eps_hat) / sqrt(ab_t) written to the last input channel
// (train.py:40-51, 404-407).  ab = alpha_bar table (NOT clamped, as in train.py:49).
__global__ void __launch_bounds__(256) selfcond_kernel(float* __restrict__ net, const float* __restrict__ eps_hat,
                                                       const int64_t* __restrict__ t, const float* __restrict__ ab, int Cx, int L) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.y;
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const float a = ab[t[b]];
    float* nb = net + (size_t)b * Cx * L;
    nb[(size_t)(Cx - 1) * L + l] = __fdiv_rn(__fsub_rn(nb[l], __fmul_rn(sqrtf(1.0f - a), eps_hat[(size_t)b * L + l])), sqrtf(a));
}
extern "C" int gw_selfcond_x0(float* net, const float* eps_hat, const int64_t* t, const float* alpha_bar, int B, int Cx, int L,
                              void* stream) {
    GW_REQUIRE(B > 0 && L > 0 && Cx >= 2, "gw_selfcond_x0: sizes");
    GW_CUDA(gw_launch_pdl(selfcond_kernel, dim3(gw_cdiv(L, 256), B), dim3(256), (size_t)(0), (cudaStream_t)stream, net, eps_hat, t, alpha_bar, Cx, L));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------
// time-MLP / tproj backward.  aux rows (written by gw_film_vectors): [emb(time_dim) | pre(base) | ctx(base) | act(base)]
//   film = W2 act + b2, act = silu(ctx), ctx = silu(pre), pre = W1 emb + b1
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dsilu(float x) {
    const float s = 1.0f / (1.0f + expf(-x));
    return s * (1.0f + x * (1.0f - s));
}
// grid F/4, block (64, 4, BL): thread (jl, f, bl) sums the batch lane bl, bl+BL, ... for columns j = jl, jl+64, ...; the
// lanes are folded in fixed order
#define FILM_BL 4
__global__ void __launch_bounds__(1024) film_bwd_w2_kernel(const float* __restrict__ dfilm, const float* __restrict__ aux, int B,
                                                           int td, int base, int F, float* __restrict__ dW2,
                                                           float* __restrict__ db2) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ float red[];                   // [BL][4][base]
    __shared__ float redb[FILM_BL][4];
    const int jl = threadIdx.x, fl = threadIdx.y, bl = threadIdx.z;
    const int f = blockIdx.x * 4 + fl;
    const int na = td + 3 * base;
    float bsum = 0.0f;
    if (f < F && jl == 0) {
        for (int b = bl; b < B; b += FILM_BL) bsum += dfilm[(size_t)b * F + f];
    }
    for (int j = jl; j < base; j += 64) {
        float acc = 0.0f;
        if (f < F) {
#pragma unroll 8
            for (int b = bl; b < B; b += FILM_BL)
                acc = fmaf(dfilm[(size_t)b * F + f], aux[(size_t)b * na + td + 2 * base + j], acc);
        }
        red[(bl * 4 + fl) * base + j] = acc;
    }
    if (jl == 0) redb[bl][fl] = bsum;
    __syncthreads();
    if (bl == 0 && f < F) {
        for (int j = jl; j < base; j += 64) {
            float a = 0.0f;
#pragma unroll
            for (int t = 0; t < FILM_BL; ++t) a += red[(t * 4 + fl) * base + j];
            dW2[(size_t)f * base + j] += a;
        }
        if (jl == 0) {
            float bb = 0.0f;
#pragma unroll
            for (int t = 0; t < FILM_BL; ++t) bb += redb[t][fl];
            db2[f] += bb;
        }
    }
}
// grid B, block 1024: dpre[b, j] = (sum_f dfilm[b,f] W2[f,j]) * silu'(ctx) * silu'(pre)
__global__ void __launch_bounds__(1024) film_bwd_act_kernel(const float* __restrict__ dfilm, const float* __restrict__ aux,
                                                            const float* __restrict__ w2, int td, int base, int F,
                                                            float* __restrict__ dpre) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ float sm[];           // [n_part][base]
    const int b = blockIdx.x;
    const int j = threadIdx.x % base, part = threadIdx.x / base, n_part = blockDim.x / base;
    float acc = 0.0f;
#pragma unroll 8
    for (int f = part; f < F; f += n_part) acc = fmaf(dfilm[(size_t)b * F + f], w2[(size_t)f * base + j], acc);
    sm[part * base + j] = acc;
    __syncthreads();
    if (part == 0) {
        float s = 0.0f;
        for (int p = 0; p < n_part; ++p) s += sm[p * base + j];
        const float* ax = aux + (size_t)b * (td + 3 * base) + td;
        dpre[(size_t)b * base + j] = s * dsilu(ax[base + j]) * dsilu(ax[j]);
    }
}
// grid base, block (128, 8): dW1[j, i] += sum_b dpre[b,j] emb[b,i] (i = il, il+128, ...); db1[j] += sum_b dpre[b,j]
__global__ void __launch_bounds__(1024) film_bwd_w1_kernel(const float* __restrict__ dpre, const float* __restrict__ aux, int B,
                                                           int td, int base, float* __restrict__ dW1, float* __restrict__ db1) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ float sm[];           // [8][td] + [8]
    const int j = blockIdx.x, il = threadIdx.x, bl = threadIdx.y;
    const int na = td + 3 * base;
    float bs = 0.0f;
    if (il == 0)
        for (int b = bl; b < B; b += 8) bs += dpre[(size_t)b * base + j];
    for (int i = il; i < td; i += 128) {
        float acc = 0.0f;
#pragma unroll 8
        for (int b = bl; b < B; b += 8) acc = fmaf(dpre[(size_t)b * base + j], aux[(size_t)b * na + i], acc);
        sm[bl * td + i] = acc;
    }
    if (il == 0) sm[8 * td + bl] = bs;
    __syncthreads();
    if (bl == 0) {
        for (int i = il; i < td; i += 128) {
            float a = 0.0f;
#pragma unroll
            for (int t = 0; t < 8; ++t) a += sm[t * td + i];
            dW1[(size_t)j * td + i] += a;
        }
        if (il == 0) {
            float bb = 0.0f;
#pragma unroll
            for (int t = 0; t < 8; ++t) bb += sm[8 * td + t];
            db1[j] += bb;
        }
    }
}
// ---- tiled versions for base in {64, 128} (the strided per-thread batch loops above are latency-bound: 95 us per step at
// B = 256 for 75 MFLOP).  Summation orders are fixed, so results are reproducible run to run.
// dW2[f, j] += sum_b dfilm[b, f] act[b, j], db2[f] += sum_b dfilm[b, f]: CTA = 32 rows f x all j, batch tiles of 32 in shared memory
template <int BASE>
__global__ void __launch_bounds__(256) film_bwd_w2_tiled_kernel(const float* __restrict__ dfilm, const float* __restrict__ aux, int B,
                                                                int td, int F, float* __restrict__ dW2, float* __restrict__ db2) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int FG = 256 / BASE, NF = 32 / FG;
    __shared__ float df_s[32][33];
    __shared__ float act_s[32][BASE];
    const int f0 = blockIdx.x * 32, tid = threadIdx.x;
    const int j = tid % BASE, fg = tid / BASE;
    const int na = td + 3 * BASE;
    float acc[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) acc[i] = 0.0f;
    float bs = 0.0f;
    for (int b0 = 0; b0 < B; b0 += 32) {
        for (int i = tid; i < 32 * 32; i += 256) {
            const int bb = i >> 5, ff = i & 31;
            df_s[bb][ff] = (b0 + bb < B && f0 + ff < F) ? dfilm[(size_t)(b0 + bb) * F + f0 + ff] : 0.0f;
        }
        for (int i = tid; i < 32 * BASE; i += 256) {
            const int bb = i / BASE, jj = i % BASE;
            act_s[bb][jj] = b0 + bb < B ? aux[(size_t)(b0 + bb) * na + td + 2 * BASE + jj] : 0.0f;
        }
        __syncthreads();
#pragma unroll 8
        for (int bb = 0; bb < 32; ++bb) {
            const float a = act_s[bb][j];
#pragma unroll
            for (int i = 0; i < NF; ++i) acc[i] = fmaf(df_s[bb][fg * NF + i], a, acc[i]);
        }
        if (tid < 32) {
#pragma unroll 8
            for (int bb = 0; bb < 32; ++bb) bs += df_s[bb][tid];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < NF; ++i) {
        const int f = f0 + fg * NF + i;
        if (f < F) dW2[(size_t)f * BASE + j] += acc[i];
    }
    if (tid < 32 && f0 + tid < F) db2[f0 + tid] += bs;
}
// partial[ks, b, j] = sum_{f in split ks} dfilm[b, f] W2[f, j]: CTA = 8 samples x one of KS f-ranges (FS = 96 values of f)
#define FILM_FS 96
template <int BASE>
__global__ void __launch_bounds__(256) film_bwd_act_split_kernel(const float* __restrict__ dfilm, const float* __restrict__ w2, int B,
                                                                 int F, float* __restrict__ partial) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int SG = 256 / BASE, NS = 8 / SG;          // sample groups, samples per thread
    __shared__ float df_s[8][FILM_FS];
    const int b0 = blockIdx.x * 8, ks = blockIdx.y, fs0 = ks * FILM_FS, tid = threadIdx.x;
    const int j = tid % BASE, sg = tid / BASE;
    for (int i = tid; i < 8 * FILM_FS; i += 256) {
        const int s = i / FILM_FS, ff = i % FILM_FS;
        df_s[s][ff] = (b0 + s < B && fs0 + ff < F) ? dfilm[(size_t)(b0 + s) * F + fs0 + ff] : 0.0f;
    }
    __syncthreads();
    float acc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s] = 0.0f;
    const int nf = min(FILM_FS, F - fs0);
#pragma unroll 16
    for (int ff = 0; ff < nf; ++ff) {
        const float w = w2[(size_t)(fs0 + ff) * BASE + j];
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s] = fmaf(df_s[sg * NS + s][ff], w, acc[s]);
    }
#pragma unroll
    for (int s = 0; s < NS; ++s)
        if (b0 + sg * NS + s < B) partial[((size_t)ks * B + b0 + sg * NS + s) * BASE + j] = acc[s];
}
// dpre[b, j] = (sum_ks partial[ks, b, j]) * silu'(ctx) * silu'(pre)
__global__ void film_bwd_act_fold_kernel(const float* __restrict__ partial, int n_split, int B, const float* __restrict__ aux,
                                         int td, int base, float* __restrict__ dpre) {
    pdl_wait();
    pdl_launch_dependents();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * base) return;
    const int b = i / base, j = i % base;
    float s = 0.0f;
    for (int k = 0; k < n_split; ++k) s += partial[(size_t)k * B * base + i];
    const float* ax = aux + (size_t)b * (td + 3 * base) + td;
    dpre[i] = s * dsilu(ax[base + j]) * dsilu(ax[j]);
}

extern "C" int gw_film_bwd(const float* dfilm, const float* aux, const float* w2, int B, int time_dim, int base, int F,
                           float* scratch, float* dW1, float* db1, float* dW2, float* db2, void* stream) {
    GW_REQUIRE(B > 0 && base > 0 && base <= 1024 && 1024 % base == 0 && time_dim > 0 && time_dim <= 4096,
               "gw_film_bwd: sizes (base must divide 1024)");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_split = gw_cdiv(F, FILM_FS);
    if (base == 64 || base == 128) {
        float* partial = scratch + (size_t)B * base;               // [n_split, B, base] behind dpre [B, base]
        if (base == 64) {
            GW_CUDA(gw_launch_pdl(film_bwd_w2_tiled_kernel<64>, dim3(gw_cdiv(F, 32)), dim3(256), (size_t)(0), st, dfilm, aux, B, time_dim, F, dW2, db2));
            GW_CUDA(gw_launch_pdl(film_bwd_act_split_kernel<64>, dim3(gw_cdiv(B, 8), n_split), dim3(256), (size_t)(0), st, dfilm, w2, B, F, partial));
        } else {
            GW_CUDA(gw_launch_pdl(film_bwd_w2_tiled_kernel<128>, dim3(gw_cdiv(F, 32)), dim3(256), (size_t)(0), st, dfilm, aux, B, time_dim, F, dW2, db2));
            GW_CUDA(gw_launch_pdl(film_bwd_act_split_kernel<128>, dim3(gw_cdiv(B, 8), n_split), dim3(256), (size_t)(0), st, dfilm, w2, B, F, partial));
        }
        GW_LAUNCH_CHECK();
        GW_CUDA(gw_launch_pdl(film_bwd_act_fold_kernel, dim3(gw_cdiv(B * base, 256)), dim3(256), (size_t)(0), st, partial, n_split, B, aux, time_dim, base, scratch));
        GW_LAUNCH_CHECK();
        GW_CUDA(gw_launch_pdl(film_bwd_w1_kernel, dim3(base), dim3(128, 8), (size_t)((size_t)(8 * time_dim + 8) * sizeof(float)), st, scratch, aux, B, time_dim, base, dW1, db1));
        GW_LAUNCH_CHECK();
        return GW_OK;
    }
    GW_CUDA(gw_launch_pdl(film_bwd_w2_kernel, dim3(gw_cdiv(F, 4)), dim3(64, 4, FILM_BL), (size_t)((size_t)FILM_BL * 4 * base * sizeof(float)), st, dfilm, aux, B, time_dim,
                                                                                                          base, F, dW2, db2));
    GW_LAUNCH_CHECK();
    GW_CUDA(gw_launch_pdl(film_bwd_act_kernel, dim3(B), dim3(1024), (size_t)((size_t)1024 * sizeof(float)), st, dfilm, aux, w2, time_dim, base, F, scratch));
    GW_LAUNCH_CHECK();
    GW_CUDA(gw_launch_pdl(film_bwd_w1_kernel, dim3(base), dim3(128, 8), (size_t)((size_t)(8 * time_dim + 8) * sizeof(float)), st, scratch, aux, B, time_dim, base, dW1, db1));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------
// clip_grad_norm_ + AdamW + EMA over flat buffers
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long n, double* __restrict__ partial) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ double red[8];
    double a = 0.0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double v = (double)g[i];
        a += v * v;
    }
    a = warp_sum_d(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i];
        partial[blockIdx.x] = s;
    }
}

// hyper (device, fp32[16]) -- constants of the run, uploaded once (nothing here changes per step):
//   0 base lr, 3 ema_decay (<0: no EMA), 4 weight_decay, 5 max_norm (<=0: no clipping), 6 grad_scale (1/world, applied before the
//   norm), 7 skip_loss_threshold (<=0: off; train.py:428-436), 8 warmup_steps, 9 total_steps, 10 min_lr_scale, 11 use_sched (0/1)
// state (device, int32[4]): 0 = optimisation steps APPLIED so far (train.py: a skipped batch `continue`s before optimizer.step
//   and scheduler.step, so neither the bias correction nor the LR schedule advances); advanced by gw_train_advance.
// info (device, fp32[8]) out: 0 total grad norm (after grad_scale), 1 clip coefficient, 2 1 if the step was applied, 3 lr used,
//   4 batch loss (mean over ranks when the bucket carries it)
// lossv: the batch loss; with g_has_loss it is g[n] * grad_scale (the loss rides in the bucket's extra slot through the
//   all-reduce, so every rank takes the same skip decision), else *loss (may be NULL).
__device__ __forceinline__ double warmup_cosine_dev(double step, double warmup, double total, double min_scale) {   // train.py:85-90
    if (step < warmup) return fmax(1e-8, (step + 1.0) / fmax(1.0, warmup));
    double progress = (step - warmup) / fmax(1.0, total - warmup);
    progress = fmin(fmax(progress, 0.0), 1.0);
    return min_scale + 0.5 * (1.0 - min_scale) * (1.0 + cos(3.14159265358979323846 * progress));
}
__global__ void __launch_bounds__(256) adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, float* __restrict__ ema, long n,
                                                        const double* __restrict__ partial, int n_partial,
                                                        const float* __restrict__ hyper, const float* __restrict__ loss,
                                                        int g_has_loss, const int* __restrict__ state, double beta1d,
                                                        double beta2d, float eps, float* __restrict__ info) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float s_coef, s_lr, s_bc1, s_bc2s;
    __shared__ int s_ok;
    if (threadIdx.x < 32) {
        double a = 0.0;
        for (int i = threadIdx.x; i < n_partial; i += 32) a += partial[i];
        a = warp_sum_d(a);
        if (threadIdx.x == 0) {
            const float gs = hyper[6];
            const float norm = (float)sqrt(a) * gs;
            const float max_norm = hyper[5];
            float coef = 1.0f;
            if (max_norm > 0.0f) coef = fminf(max_norm / (norm + 1e-6f), 1.0f);    // torch clip_grad_norm_
            float lossv = 0.0f;
            bool have_loss = false;
            if (g_has_loss) { lossv = g[n] * gs; have_loss = true; }
            else if (loss != nullptr) { lossv = loss[0]; have_loss = true; }
            const float thr = hyper[7];
            bool ok = isfinite(norm) && (!have_loss || isfinite(lossv));            // train.py:424-427 skips the batch
            if (ok && have_loss && thr > 0.0f && lossv > thr) ok = false;           // train.py:428-436 (--skip_bad_batches)
            const int applied = state != nullptr ? state[0] : 0;
            const double lam = hyper[11] != 0.0f ? warmup_cosine_dev((double)applied, (double)hyper[8], (double)hyper[9], (double)hyper[10]) : 1.0;
            const float lr = (float)((double)hyper[0] * lam);
            const double stepd = (double)(applied + 1);
            s_coef = coef * gs;
            s_ok = ok ? 1 : 0;
            s_lr = lr;
            s_bc1 = (float)(1.0 - pow(beta1d, stepd));
            s_bc2s = (float)sqrt(1.0 - pow(beta2d, stepd));
            if (blockIdx.x == 0) {
                info[0] = norm;
                info[1] = coef;
                info[2] = ok ? 1.0f : 0.0f;
                info[3] = lr;
                info[4] = lossv;
            }
        }
    }
    __syncthreads();
    if (!s_ok) return;
    const float coef = s_coef;
    const float lr = s_lr, bc1 = s_bc1, bc2s = s_bc2s, decay = hyper[3], wd = hyper[4];
    const float beta1 = (float)beta1d, beta2 = (float)beta2d;
    const float step_size = lr / bc1;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        float pi = p[i] * (1.0f - lr * wd);                        // torch.optim.AdamW (decoupled decay first)
        const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);      // exp_avg.lerp_(grad, 1-beta1)
        const float vi = v[i] * beta2 + (1.0f - beta2) * gi * gi;
        const float denom = sqrtf(vi) / bc2s + eps;
        pi -= step_size * (mi / denom);
        p[i] = pi;
        m[i] = mi;
        v[i] = vi;
        if (ema != nullptr && decay >= 0.0f) ema[i] = ema[i] * decay + pi * (1.0f - decay);     // train.py:78
    }
}

#define OPT_PARTIALS 296
extern "C" int gw_opt_scratch_doubles(void) { return OPT_PARTIALS; }
extern "C" int gw_grad_sumsq(const float* g, long n, double* partial, void* stream) {
    GW_REQUIRE(n > 0, "gw_grad_sumsq: n");
    GW_CUDA(gw_launch_pdl(sumsq_kernel, dim3(OPT_PARTIALS), dim3(256), (size_t)(0), (cudaStream_t)stream, g, n, partial));
    GW_LAUNCH_CHECK();
    return GW_OK;
}
extern "C" int gw_adamw_ema(float* p, const float* g, float* m, float* v, float* ema, long n, const double* partial,
                            const float* hyper, const float* loss, int g_has_loss, const int* state, double beta1, double beta2,
                            float eps, float* info, void* stream) {
    GW_REQUIRE(n > 0 && hyper != nullptr && info != nullptr && partial != nullptr, "gw_adamw_ema: arguments");
    int grid = (int)((n + 1023) / 1024);
    if (grid > 148 * 4) grid = 148 * 4;
    GW_CUDA(gw_launch_pdl(adamw_ema_kernel, grid, dim3(256), (size_t)(0), (cudaStream_t)stream, p, g, m, v, ema, n, partial, OPT_PARTIALS, hyper, loss, g_has_loss, state,
                                                             beta1, beta2, eps, info));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// gradient bucket reset before the backward pass: g[0, n) = 0 and (loss != NULL) g[n] = *loss, the extra slot that carries the
// batch loss through the all-reduce
__global__ void __launch_bounds__(256) bucket_reset_kernel(float4* __restrict__ g4, long n4, float* __restrict__ g, long n,
                                                           const float* __restrict__ loss) {
    pdl_wait();
    pdl_launch_dependents();
    const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) g4[i] = z;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (long i = n4 * 4; i < n; ++i) g[i] = 0.0f;
        if (loss != nullptr) g[n] = loss[0];
    }
}
extern "C" int gw_bucket_reset(float* g, long n, const float* loss, void* stream) {
    GW_REQUIRE(g != nullptr && n > 0 && ((uintptr_t)g & 15) == 0, "gw_bucket_reset: g must be 16-byte aligned");
    int grid = (int)((n / 4 + 255) / 256);
    if (grid > 148 * 2) grid = 148 * 2;
    if (grid < 1) grid = 1;
    GW_CUDA(gw_launch_pdl(bucket_reset_kernel, grid, dim3(256), (size_t)(0), (cudaStream_t)stream, reinterpret_cast<float4*>(g), n / 4, g, n, loss));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// end of an optimisation step: the Philox draw counter always advances (the reference consumes its RNG on skipped batches
// too), the applied-step counter only when gw_adamw_ema applied the update (info[2])
__global__ void train_advance_kernel(int* step_ctr, int* state, const float* info) {
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (step_ctr != nullptr) *step_ctr += 1;
        if (state != nullptr) {
            if (info == nullptr || info[2] != 0.0f) state[0] += 1;
            else state[1] += 1;                                 // skipped batches (train.py: skipped_batches)
        }
    }
}
extern "C" int gw_train_advance(int* step_ctr, int* state, const float* info, void* stream) {
    GW_CUDA(gw_launch_pdl(train_advance_kernel, dim3(1), dim3(32), (size_t)(0), (cudaStream_t)stream, step_ctr, state, info));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// loss weight (train.py:414-417): wt[b] = (1 - alpha_bar[t_b])^power
__global__ void loss_weight_kernel(const long long* __restrict__ t, const float* __restrict__ ab, float power, float* __restrict__ wt, int B) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) wt[b] = powf(1.0f - ab[t[b]], power);
}
extern "C" int gw_loss_weight(const int64_t* t, const float* alpha_bar, float power, float* wt, int B, void* stream) {
    GW_REQUIRE(t != nullptr && alpha_bar != nullptr && wt != nullptr && B > 0, "gw_loss_weight: arguments");
    GW_CUDA(gw_launch_pdl(loss_weight_kernel, dim3(gw_cdiv(B, 128)), dim3(128), (size_t)(0), (cudaStream_t)stream, (const long long*)t, alpha_bar, power, wt, B));
    GW_LAUNCH_CHECK();
    return GW_OK;
}
