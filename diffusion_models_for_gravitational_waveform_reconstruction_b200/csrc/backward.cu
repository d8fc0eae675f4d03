// Backward-path kernels of the training step (train.py:408-439 of the reference = autograd through
// UNet1D.forward + the masked Huber/MSE loss).  Everything here is CUDA-core fp32 math on channels-last
// activations; the tcgen05 dgrad / wgrad GEMMs live in conv_tc.cu / wgrad_tc.cu.
//
// Notation for one conv block (models.py:160-173, 188-193), per sample b, position l, channel c (group g):
//   z   = conv(in) + bias                       ("raw", saved by the forward)
//   xh  = (z - mean[b,g]) * rstd[b,g]
//   n   = xh * gn_w[c] + gn_b[c]
//   a   = n * sigmoid(n)
//   h   = a + bc[c] + sum_j wc[c,j] * cond[b,l,j]
//   o   = h * (1 + gamma[b,c]) + beta[b,c]      ("out"; encoders also emit pooled = avg of row pairs)
// Given do = dL/do:  dbeta = sum_l do, dgamma = sum_l do*h, dh = do*(1+gamma), dn = dh * silu'(n),
//   d gn_w = sum dn*xh, d gn_b = sum dn, dxh = dn*gn_w,
//   dz = rstd * (dxh - mean_g(dxh) - xh * mean_g(dxh*xh))      (biased variance, models.py:154-158)
// All reductions are two-level and deterministic: per-CTA partials in caller-provided scratch, then a
// fixed-order second pass (no floating-point atomics anywhere).
#include "common.cuh"
#include "../../include/gwb200.h"
#include <string.h>

#define BW_MAX_CC 8

// ------------------------------------------------------------------------------------------------
// generic column sums: dst[c] (+)= scale * sum_r src[r, c], fixed summation order (deterministic).
// Block = 32 columns x 32 row lanes.  Tall inputs are reduced in two stages; stage 1 folds each chunk of 128 rows IN PLACE
// into the chunk's first row (the partial buffers are scratch), stage 2 folds those rows into dst.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) reduce_rows_kernel(float* __restrict__ src, int n_rows, long n_cols, long pitch, int chunk,
                                                           float scale, float* __restrict__ dst, int accumulate, int final_stage) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float red[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long c = (long)blockIdx.x * 32 + tx;
    const int r0 = blockIdx.y * chunk;
    const int r1 = min(r0 + chunk, n_rows);
    float a = 0.0f;
    if (c < n_cols) {
#pragma unroll 4
        for (int r = r0 + ty; r < r1; r += 32) a += src[(size_t)r * pitch + c];
    }
    red[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && c < n_cols) {
        float s = 0.0f;
#pragma unroll
        for (int t = 0; t < 32; ++t) s += red[t][tx];
        if (final_stage) dst[c] = accumulate ? dst[c] + scale * s : scale * s;
        else src[(size_t)r0 * pitch + c] = s;
    }
}

static int reduce_rows(const float* src_c, int n_rows, long n_cols, long pitch, float scale, float* dst, int accumulate,
                       cudaStream_t st) {
    float* src = const_cast<float*>(src_c);
    const unsigned gx = (unsigned)((n_cols + 31) / 32);
    if (n_rows <= 256) {
        GW_CUDA(gw_launch_pdl(reduce_rows_kernel, dim3(gx, 1), dim3(1024), (size_t)(0), st, src, n_rows, n_cols, pitch, n_rows, scale, dst, accumulate, 1));
        GW_LAUNCH_CHECK();
        return GW_OK;
    }
    const int chunk = 128;
    const int n_chunks = gw_cdiv(n_rows, chunk);
    GW_CUDA(gw_launch_pdl(reduce_rows_kernel, dim3(gx, n_chunks), dim3(1024), (size_t)(0), st, src, n_rows, n_cols, pitch, chunk, 1.0f, nullptr, 0, 0));
    GW_LAUNCH_CHECK();
    if (n_chunks <= 256) {
        GW_CUDA(gw_launch_pdl(reduce_rows_kernel, dim3(gx, 1), dim3(1024), (size_t)(0), st, src, n_chunks, n_cols, pitch * chunk, n_chunks, scale, dst, accumulate, 1));
        GW_LAUNCH_CHECK();
        return GW_OK;
    }
    return reduce_rows(src, n_chunks, n_cols, pitch * chunk, scale, dst, accumulate, st);
}

extern "C" int gw_reduce_rows(const float* src, int n_rows, long n_cols, float scale, float* dst, int accumulate, void* stream) {
    GW_REQUIRE(n_rows > 0 && n_cols > 0, "gw_reduce_rows: sizes");
    return reduce_rows(src, n_rows, n_cols, n_cols, scale, dst, accumulate, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// loss forward + dL/d eps_hat  (train.py:53-58, 411-421)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
}

__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ eps_hat, const float* __restrict__ eps,
                                                   const float* __restrict__ mask, const float* __restrict__ wt, int B, int L,
                                                   int loss_type, float beta, float grad_scale, float* __restrict__ per_sample,
                                                   float* __restrict__ d_eps) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float red[8];
    const int b = blockIdx.x;
    const float* eh = eps_hat + (size_t)b * L;
    const float* e = eps + (size_t)b * L;
    const float* m = mask ? mask + (size_t)b * L : nullptr;
    float sm = 0.0f, sl = 0.0f;
    for (int l = threadIdx.x; l < L; l += 256) {
        const float mk = m ? m[l] : 1.0f;
        const float d = eh[l] - e[l];
        float el;
        if (loss_type == 0) {
            const float ad = fabsf(d);
            el = ad < beta ? 0.5f * d * d / beta : ad - 0.5f * beta;      // F.smooth_l1_loss(beta)
        } else {
            el = d * d;
        }
        sm += mk;
        sl += el * mk;
    }
    const float tm = block_sum_256(sm, red);
    const float tl = block_sum_256(sl, red);
    const float denom = fmaxf(tm, 1.0f);                                    // mask.sum().clamp_min(1) (train.py:419)
    const float w = wt ? wt[b] : 1.0f;
    if (threadIdx.x == 0) per_sample[b] = w * tl / denom;
    const float k = grad_scale * w / (denom * (float)B);
    for (int l = threadIdx.x; l < L; l += 256) {
        const float mk = m ? m[l] : 1.0f;
        const float d = eh[l] - e[l];
        float g;
        if (loss_type == 0) g = fabsf(d) < beta ? d / beta : (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f));
        else g = 2.0f * d;
        d_eps[(size_t)b * L + l] = k * mk * g;
    }
}

__global__ void __launch_bounds__(256) mean_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float red[8];
    float a = 0.0f;
    for (int i = threadIdx.x; i < n; i += 256) a += v[i];
    const float t = block_sum_256(a, red);
    if (threadIdx.x == 0) out[0] = t / (float)n;
}

extern "C" int gw_loss(const float* eps_hat, const float* eps, const float* mask, const float* wt, int B, int L, int loss_type,
                       float beta, float grad_scale, float* per_sample, float* loss, float* d_eps, void* stream) {
    GW_REQUIRE(B > 0 && L > 0 && (loss_type == 0 || loss_type == 1), "gw_loss: arguments");
    GW_REQUIRE(loss_type == 1 || beta > 0.0f, "gw_loss: huber beta must be > 0");
    GW_CUDA(gw_launch_pdl(loss_kernel, dim3(B), dim3(256), (size_t)(0), (cudaStream_t)stream, eps_hat, eps, mask, wt, B, L, loss_type, beta, grad_scale, per_sample, d_eps));
    GW_LAUNCH_CHECK();
    GW_CUDA(gw_launch_pdl(mean_kernel, dim3(1), dim3(256), (size_t)(0), (cudaStream_t)stream, per_sample, B, loss));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------
// head conv backward (models.py:230): d_h, and per-CTA partials of d final.weight / d final.bias
//   eps[l] = bias + sum_k sum_c wf[c,k] * hcat[l+k-1, c]   =>   d_hcat[l,c] = sum_k wf[c,k] * d_eps[l-k+1]
// partial layout per CTA: [(C+1)*3 + 1]  (weight grads in the reference's [c][k] order, then the bias grad)
// ------------------------------------------------------------------------------------------------
template <typename T, bool WRITE_DH>
__global__ void __launch_bounds__(256) final_bwd_kernel(const float* __restrict__ d_eps, const T* __restrict__ h,
                                                        const float* __restrict__ net, int Cx, int L, int C,
                                                        const float* __restrict__ wf, T* __restrict__ d_h,
                                                        float* __restrict__ partial, int rows_per_cta) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ float sm[];
    const int n_oct = C / 8;
    const int n_tr = 256 / n_oct;                 // thread rows
    float* red = sm;                              // [n_tr][C*3]
    const int b = blockIdx.y, r0 = blockIdx.x * rows_per_cta;
    const int oct = threadIdx.x % n_oct, tr = threadIdx.x / n_oct;
    float w[8][3], dw[8][3];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            w[i][k] = WRITE_DH ? wf[(oct * 8 + i) * 3 + k] : 0.0f;
            dw[i][k] = 0.0f;
        }
    const float* de = d_eps + (size_t)b * L;
    const int r_end = min(r0 + rows_per_cta, L);
    constexpr int UN = 4;
    for (int r = r0 + tr; r < r_end; r += n_tr * UN) {
        float hv[UN][8], em[UN], ec[UN], ep[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int rr = r + u * n_tr;
            const int rc = rr < r_end ? rr : r;
            ld8(h + ((size_t)b * L + rc) * C + oct * 8, hv[u]);
            em[u] = rc > 0 ? de[rc - 1] : 0.0f;
            ec[u] = de[rc];
            ep[u] = rc + 1 < L ? de[rc + 1] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int rr = r + u * n_tr;
            if (rr >= r_end) break;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (WRITE_DH) o[i] = fmaf(w[i][0], ep[u], fmaf(w[i][1], ec[u], w[i][2] * em[u]));
                dw[i][0] = fmaf(hv[u][i], ep[u], dw[i][0]);
                dw[i][1] = fmaf(hv[u][i], ec[u], dw[i][1]);
                dw[i][2] = fmaf(hv[u][i], em[u], dw[i][2]);
            }
            if (WRITE_DH) st8(d_h + ((size_t)b * L + rr) * C + oct * 8, o);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) red[(size_t)tr * C * 3 + (oct * 8 + i) * 3 + k] = dw[i][k];
    __syncthreads();
    float* pt = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * ((C + 1) * 3 + 1);
    for (int i = threadIdx.x; i < C * 3; i += 256) {
        float a = 0.0f;
        for (int t = 0; t < n_tr; ++t) a += red[(size_t)t * C * 3 + i];
        pt[i] = a;
    }
    // x_t channel (index C of hcat) and the bias: warp 0
    if (threadIdx.x < 32) {
        const float* xr = net + (size_t)b * Cx * L;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, ab = 0.0f;
        for (int r = r0 + threadIdx.x; r < r_end; r += 32) {
            const float e_m = r > 0 ? de[r - 1] : 0.0f, e_c = de[r], e_p = r + 1 < L ? de[r + 1] : 0.0f;
            const float x = xr[r];
            a0 = fmaf(x, e_p, a0);
            a1 = fmaf(x, e_c, a1);
            a2 = fmaf(x, e_m, a2);
            ab += e_c;
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); ab = warp_sum(ab);
        if (threadIdx.x == 0) {
            pt[C * 3 + 0] = a0; pt[C * 3 + 1] = a1; pt[C * 3 + 2] = a2; pt[C * 3 + 3] = ab;
        }
    }
}

int final_bwd_stream(const float* d_eps, const void* h, const float* net, int B, int Cx, int L, const float* wf, void* d_h,
                     float* partial, int* n_cta, cudaStream_t st);
int g_final_bwd_stream = 1;
extern "C" int gw_final_bwd(const float* d_eps, const void* h, int dtype, const float* net, int B, int Cx, int L, int C,
                            const float* wf, void* d_h, float* scratch, float* d_wf, float* d_bf, void* stream) {
    GW_REQUIRE(C % 64 == 0 && C <= 256, "gw_final_bwd: C=%d", C);
    GW_REQUIRE(dtype == GW_F32 || dtype == GW_BF16, "gw_final_bwd: dtype %d", dtype);
    const int rows = 1024;
    dim3 grid(gw_cdiv(L, rows), B);
    const int n_tr = 256 / (C / 8);
    const size_t smem = (size_t)n_tr * C * 3 * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GW_BF16 && C == 64 && L % 4 == 0 && g_final_bwd_stream) {       // HBM-streaming kernel (stream_gn.cu)
        int n_ctas = 0;
        int rcs = final_bwd_stream(d_eps, h, net, B, Cx, L, wf, d_h, scratch, &n_ctas, st);
        if (rcs != GW_OK) return rcs;
        const int nvs = (C + 1) * 3 + 1;
        rcs = reduce_rows(scratch, n_ctas, nvs - 1, nvs, 1.0f, d_wf, 1, st);
        if (rcs != GW_OK) return rcs;
        return reduce_rows(scratch + nvs - 1, n_ctas, 1, nvs, 1.0f, d_bf, 1, st);
    }
#define FB_GO(TT, WR)                                                                                                  \
    do {                                                                                                               \
        GW_CUDA(cudaFuncSetAttribute(final_bwd_kernel<TT, WR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        GW_CUDA(gw_launch_pdl(final_bwd_kernel<TT, WR>, grid, dim3(256), (size_t)(smem), st, d_eps, (const TT*)h, net, Cx, L, C, wf, (TT*)d_h, scratch, rows)); \
    } while (0)
    if (dtype == GW_F32) {
        if (d_h) FB_GO(float, true); else FB_GO(float, false);
    } else {
        if (d_h) FB_GO(bf16, true); else FB_GO(bf16, false);
    }
#undef FB_GO
    GW_LAUNCH_CHECK();
    const int n_cta = grid.x * grid.y, nv = (C + 1) * 3 + 1;
    int rc = reduce_rows(scratch, n_cta, nv - 1, nv, 1.0f, d_wf, 1, st);
    if (rc != GW_OK) return rc;
    return reduce_rows(scratch + nv - 1, n_cta, 1, nv, 1.0f, d_bf, 1, st);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm/SiLU/cond/FiLM backward.  Thread mapping as in gn_apply_kernel (forward.cu): a thread owns one
// channel quad for the whole CTA and walks rows r0+tr, r0+tr+n_tr, ...
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld4f(const float* p, float (&v)[4]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void ld4f(const bf16* p, float (&v)[4]) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
}
__device__ __forceinline__ void st4f(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4f(bf16* p, const float (&v)[4]) {
    uint2 r;
    r.x = pack_bf16x2(v[0], v[1]);
    r.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = r;
}

// sigmoid for the backward kernels: exact expf in fp32 mode, one MUFU (tanh.approx) in bf16 mode
template <bool FAST>
__device__ __forceinline__ float sigmoid_bw(float x) {
    if (FAST) {
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
        return fmaf(0.5f, t, 0.5f);
    }
    return 1.0f / (1.0f + expf(-x));
}

#include "gn_bwd.cuh"
#include "xchg.cuh"

// per-thread constants of one channel quad
template <int NCA>
struct QuadCoef {
    float A[4], Bn[4], G[4], cC[4], cW[4][NCA], gw[4];
    float mean, rstd;
};

template <int NC, int NCA>
__device__ __forceinline__ void load_quad(const GnBwdArgs& a, int b, int quad, int Cc, QuadCoef<NCA>& q) {
    const int cg = a.C / 8;
    const int g = (quad * 4) / cg;
    q.mean = a.stats[((size_t)b * 8 + g) * 2 + 0];
    q.rstd = a.stats[((size_t)b * 8 + g) * 2 + 1];
    const float* fr = a.film + (size_t)b * a.film_b_stride + a.film_off;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = quad * 4 + i;
        q.gw[i] = a.gn_w[c];
        q.A[i] = q.rstd * q.gw[i];
        q.Bn[i] = a.gn_b[c] - q.mean * q.A[i];
        q.G[i] = 1.0f + fr[c];
        q.cC[i] = NC > 0 ? a.bc[c] : 0.0f;
#pragma unroll
        for (int j = 0; j < NCA; ++j) q.cW[i][j] = (NC > 0 && j < Cc) ? a.wc[c * Cc + j] : 0.0f;
    }
}

// number of per-(b, c) sums the stats pass produces
#define GN_NV(cc) (4 + (cc))

template <typename T, bool FAST, int CC>
__global__ void __launch_bounds__(256, 2) gn_bwd_stats_kernel(GnBwdArgs a, float* __restrict__ partial) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int NC = CC >= 0 ? CC : BW_MAX_CC;
    constexpr int NCA = NC > 0 ? NC : 1;
    constexpr int NV = 4 + NC;
    extern __shared__ float red[];                      // [n_tr][C * nvr]
    const int Cc = CC >= 0 ? CC : a.Cc;
    const int b = blockIdx.y, C = a.C, L = a.L;
    const int n_quad = C / 4, n_tr = 256 / n_quad;
    const int quad = threadIdx.x % n_quad, tr = threadIdx.x / n_quad;
    // per-thread constants: z = x*A + Bn, xh = x*rstd + xo, G = 1 + gamma
    float cA[4], cBn[4], cG[4], rstd, xo;
    {
        const int cg = C / 8, g = (quad * 4) / cg;
        const float mean = a.stats[((size_t)b * 8 + g) * 2 + 0];
        rstd = a.stats[((size_t)b * 8 + g) * 2 + 1];
        xo = -mean * rstd;
        const float* fr = a.film + (size_t)b * a.film_b_stride + a.film_off;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = quad * 4 + i;
            cA[i] = rstd * a.gn_w[c];
            cBn[i] = a.gn_b[c] - mean * cA[i];
            cG[i] = 1.0f + fr[c];
        }
    }
    // per (b, c): 0 sum do, 1 sum do*silu(n), 2 sum dn, 3 sum dn*xh, 4+j sum do*cond_j
    float acc[4][NV];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[i][v] = 0.0f;
    const T* raw = (const T*)a.raw + (size_t)b * L * C + quad * 4;
    const T* doa = a.do_a ? (const T*)a.do_a + (size_t)b * L * C + quad * 4 : nullptr;
    const int Lp = L / 2;
    const T* dop = a.do_pool ? (const T*)a.do_pool + (size_t)b * Lp * C + quad * 4 : nullptr;
    const float* cbase = a.cond + (size_t)b * L * Cc;
    const int r0 = blockIdx.x * a.rows_per_cta;
    const int r_end = min(r0 + a.rows_per_cta, L);
    constexpr int UN = 4;                               // rows in flight per thread: every load is issued before any math
    for (int r = r0 + tr; r < r_end; r += n_tr * UN) {
        float x[UN][4], d[UN][4], pl[UN][4], cv[UN][NCA];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int rr = r + u * n_tr;
            const int rc = rr < r_end ? rr : r;         // clamped address; the result of a clamped row is discarded
            ld4f(raw + (size_t)rc * C, x[u]);
            if (doa) ld4f(doa + (size_t)rc * C, d[u]);
            if (dop) ld4f(dop + (size_t)min(rc >> 1, Lp - 1) * C, pl[u]);
#pragma unroll
            for (int j = 0; j < NCA; ++j) cv[u][j] = (NC > 0 && j < Cc) ? cbase[(size_t)rc * Cc + j] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int rr = r + u * n_tr;
            if (rr >= r_end) break;
            const bool pool_ok = dop != nullptr && (rr >> 1) < Lp;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float dv = doa ? d[u][i] : 0.0f;
                if (pool_ok) dv = fmaf(0.5f, pl[u][i], dv);
                const float z = fmaf(x[u][i], cA[i], cBn[i]);
                const float sg = sigmoid_bw<FAST>(z);
                const float dn = dv * cG[i] * (sg * fmaf(z, 1.0f - sg, 1.0f));
                const float xh = fmaf(x[u][i], rstd, xo);
                acc[i][0] += dv;
                acc[i][1] = fmaf(dv, z * sg, acc[i][1]);
                acc[i][2] += dn;
                acc[i][3] = fmaf(dn, xh, acc[i][3]);
#pragma unroll
                for (int j = 0; j < NC; ++j) acc[i][4 + j] = fmaf(dv, cv[u][j], acc[i][4 + j]);
            }
        }
    }
    const int nvr = 4 + Cc;                              // values really stored per channel
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int v = 0; v < NV; ++v)
            if (v < nvr) red[((size_t)tr * C + quad * 4 + i) * nvr + v] = acc[i][v];
    __syncthreads();
    float* pt = partial + ((size_t)b * gridDim.x + blockIdx.x) * C * nvr;
    for (int i = threadIdx.x; i < C * nvr; i += 256) {
        float sacc = 0.0f;
        for (int t = 0; t < n_tr; ++t) sacc += red[(size_t)t * C * nvr + i];
        pt[i] = sacc;
    }
}

// grid B, block 1024: reduce the row-CTA partials of one sample, emit dfilm and the group means needed by the apply pass
// nvr = values per channel in `partial` / `redb` (4 + Cc)
__global__ void __launch_bounds__(1024) gn_bwd_finalize_kernel(const float* __restrict__ partial, int n_rc, int C, int nvr, int Cc, int L,
                                                               const float* __restrict__ gn_w, const float* __restrict__ wc,
                                                               const float* __restrict__ bc, float* __restrict__ redb,
                                                               float* __restrict__ dfilm, long dfilm_b_stride, int film_off,
                                                               float* __restrict__ gstat) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ float sv[];                        // [C][nvr]
    const int b = blockIdx.x;
    const float* pb = partial + (size_t)b * n_rc * C * nvr;
    for (int i = threadIdx.x; i < C * nvr; i += blockDim.x) {
        float s = 0.0f;
#pragma unroll 8
        for (int r = 0; r < n_rc; ++r) s += pb[(size_t)r * C * nvr + i];
        sv[i] = s;
        redb[(size_t)b * C * nvr + i] = s;
    }
    __syncthreads();
    float* df = dfilm + (size_t)b * dfilm_b_stride + film_off;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float dg = sv[c * nvr + 1];                       // sum do*silu(n)
        if (Cc > 0) {                                     // + sum do*(bc + sum_j wc_j cond_j): h = silu(n) + cond bias
            dg = fmaf(bc[c], sv[c * nvr + 0], dg);
            for (int j = 0; j < Cc; ++j) dg = fmaf(wc[c * Cc + j], sv[c * nvr + 4 + j], dg);
        }
        df[c] = dg;                                       // d gamma = sum do*h
        df[C + c] = sv[c * nvr + 0];                      // d beta  = sum do
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, cg = C / 8;
    if (warp < 8) {
        float s1 = 0.0f, s2 = 0.0f;
        for (int c = warp * cg + lane; c < (warp + 1) * cg; c += 32) {
            s1 = fmaf(gn_w[c], sv[c * nvr + 2], s1);
            s2 = fmaf(gn_w[c], sv[c * nvr + 3], s2);
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) {
            const float n = (float)cg * (float)L;
            gstat[((size_t)b * 8 + warp) * 2 + 0] = s1 / n;
            gstat[((size_t)b * 8 + warp) * 2 + 1] = s2 / n;
        }
    }
}

// grid ceil(C/32), block 1024 = 32 channels x 32 batch lanes: parameter gradients that sum over the batch
// bias_part != NULL: also folds the conv-bias partials of the apply pass, [B * n_rc, C] (sum over rows of d_raw per CTA), in fixed
// order -- the two reduce launches per layer that used to do it are gone (the kernel then runs AFTER the apply pass).
__global__ void __launch_bounds__(1024) gn_bwd_param_kernel(const float* __restrict__ redb, int B, int C, int nvr, int Cc,
                                                            const float* __restrict__ film, long film_b_stride, int film_off,
                                                            float* __restrict__ d_gn_w, float* __restrict__ d_gn_b,
                                                            float* __restrict__ d_wc, float* __restrict__ d_bc,
                                                            const float* __restrict__ bias_part, int n_rc,
                                                            float* __restrict__ d_conv_bias) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float red[32][32][3 + BW_MAX_CC + 1];        // slot 3 + BW_MAX_CC: the conv-bias sum (SX)
    const int cl = threadIdx.x & 31, bl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    float a[4 + BW_MAX_CC];
#pragma unroll
    for (int v = 0; v < 4 + BW_MAX_CC; ++v) a[v] = 0.0f;
    if (c < C) {
#pragma unroll 4
        for (int b = bl; b < B; b += 32) {
            const float* p = redb + ((size_t)b * C + c) * nvr;
            const float G = 1.0f + film[(size_t)b * film_b_stride + film_off + c];
            a[0] += p[3];                                 // d gn_w
            a[1] += p[2];                                 // d gn_b
            a[2] = fmaf(G, p[0], a[2]);                   // d bc
#pragma unroll
            for (int j = 0; j < BW_MAX_CC; ++j)
                if (j < Cc) a[3 + j] = fmaf(G, p[4 + j], a[3 + j]);
            if (bias_part != nullptr) {
                const float* bp = bias_part + (size_t)b * n_rc * C + c;
                float sb = 0.0f;
                for (int r = 0; r < n_rc; ++r) sb += bp[(size_t)r * C];
                a[3 + BW_MAX_CC] += sb;
            }
        }
    }
#pragma unroll
    for (int v = 0; v < 4 + BW_MAX_CC; ++v) red[bl][cl][v] = a[v];
    __syncthreads();
    // thread (v = bl, channel = cl) folds the 32 batch lanes of value v in fixed order
    if (bl < 3 + Cc && c < C) {
        float sv = 0.0f;
#pragma unroll
        for (int t = 0; t < 32; ++t) sv += red[t][cl][bl];
        if (bl == 0) d_gn_w[c] += sv;
        else if (bl == 1) d_gn_b[c] += sv;
        else if (Cc > 0) {
            if (bl == 2) d_bc[c] += sv;
            else d_wc[c * Cc + (bl - 3)] += sv;
        }
    } else if (bias_part != nullptr && bl == 31 && c < C) {
        float sv = 0.0f;
#pragma unroll
        for (int t = 0; t < 32; ++t) sv += red[t][cl][3 + BW_MAX_CC];
        d_conv_bias[c] += sv;
    }
}

template <typename T, bool FAST, int CC>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(GnBwdArgs a, const float* __restrict__ gstat, T* __restrict__ d_raw,
                                                           float* __restrict__ partial_bias) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int NC = CC >= 0 ? CC : BW_MAX_CC;
    constexpr int NCA = NC > 0 ? NC : 1;
    extern __shared__ float red[];                      // [n_tr][C]
    const int Cc = CC >= 0 ? CC : a.Cc;
    const int b = blockIdx.y, C = a.C, L = a.L;
    const int n_quad = C / 4, n_tr = 256 / n_quad;
    const int quad = threadIdx.x % n_quad, tr = threadIdx.x / n_quad;
    QuadCoef<NCA> q;
    load_quad<NC, NCA>(a, b, quad, Cc, q);
    const int g = (quad * 4) / (C / 8);
    const float m1 = gstat[((size_t)b * 8 + g) * 2 + 0], m2 = gstat[((size_t)b * 8 + g) * 2 + 1];
    float sb[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const T* raw = (const T*)a.raw + (size_t)b * L * C + quad * 4;
    const T* doa = a.do_a ? (const T*)a.do_a + (size_t)b * L * C + quad * 4 : nullptr;
    const int Lp = L / 2;
    const T* dop = a.do_pool ? (const T*)a.do_pool + (size_t)b * Lp * C + quad * 4 : nullptr;
    T* out = d_raw + (size_t)b * L * C + quad * 4;
    const int r0 = blockIdx.x * a.rows_per_cta;
    const int r_end = min(r0 + a.rows_per_cta, L);
    constexpr int UN = 4;
    for (int r = r0 + tr; r < r_end; r += n_tr * UN) {
        float x[UN][4], d[UN][4], pl[UN][4];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int rr = r + u * n_tr;
            const int rc = rr < r_end ? rr : r;
            ld4f(raw + (size_t)rc * C, x[u]);
            if (doa) ld4f(doa + (size_t)rc * C, d[u]);
            if (dop) ld4f(dop + (size_t)min(rc >> 1, Lp - 1) * C, pl[u]);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int rr = r + u * n_tr;
            if (rr >= r_end) break;
            const bool pool_ok = dop != nullptr && (rr >> 1) < Lp;
            float dz[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float dv = doa ? d[u][i] : 0.0f;
                if (pool_ok) dv = fmaf(0.5f, pl[u][i], dv);
                const float z = fmaf(x[u][i], q.A[i], q.Bn[i]);
                const float s = sigmoid_bw<FAST>(z);
                const float dn = dv * q.G[i] * (s * fmaf(z, 1.0f - s, 1.0f));
                const float xh = (x[u][i] - q.mean) * q.rstd;
                const float v = q.rstd * (fmaf(dn, q.gw[i], -m1) - xh * m2);
                dz[i] = round_to(v, d_raw);
                sb[i] += dz[i];
            }
            st4f(out + (size_t)rr * C, dz);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) red[(size_t)tr * C + quad * 4 + i] = sb[i];
    __syncthreads();
    float* pt = partial_bias + ((size_t)b * gridDim.x + blockIdx.x) * C;
    for (int i = threadIdx.x; i < C; i += 256) {
        float s = 0.0f;
        for (int t = 0; t < n_tr; ++t) s += red[(size_t)t * C + i];
        pt[i] = s;
    }
}

// ---- bf16 fast paths of the two kernels above on packed fp32x2 math (same thread mapping, same partial layouts)
__device__ __forceinline__ void bf16x4_to_pairs(const bf16* p, f32x2 (&v)[2]) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    v[0] = pk2(r.x << 16, r.x & 0xffff0000u);
    v[1] = pk2(r.y << 16, r.y & 0xffff0000u);
}
// per pair: z, silu(z) and d silu/dz from one tanh.approx per element
__device__ __forceinline__ void silu_pair(f32x2 x, f32x2 hA, f32x2 hB, f32x2& z, f32x2& act, f32x2& dact) {
    const f32x2 one = pkf2(1.0f, 1.0f), half2 = pkf2(0.5f, 0.5f), neg1 = pkf2(-1.0f, -1.0f);
    const f32x2 hh = ffma2(x, hA, hB);                   // z / 2
    z = fadd2(hh, hh);
    float h0, h1, t0, t1;
    upk2(hh, h0, h1);
    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
    const f32x2 sg = ffma2(pkf2(t0, t1), half2, half2);  // sigmoid(z)
    act = fmul2(z, sg);
    dact = fmul2(sg, ffma2(z, ffma2(sg, neg1, one), one));   // sg * (1 + z (1 - sg))
}

template <int CC, bool HEAD>
__global__ void __launch_bounds__(256, 2) gn_bwd_stats_bf16_kernel(GnBwdArgs a, float* __restrict__ partial) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int NC = CC >= 0 ? CC : BW_MAX_CC;
    constexpr int NCA = NC > 0 ? NC : 1;
    constexpr int NV = 4 + NC;
    extern __shared__ float red[];                      // [n_tr][C * nvr]
    const int Cc = CC >= 0 ? CC : a.Cc;
    const int b = blockIdx.y, C = a.C, L = a.L;
    const int n_quad = C / 4, n_tr = 256 / n_quad;
    const int quad = threadIdx.x % n_quad, tr = threadIdx.x / n_quad;
    f32x2 hA[2], hB[2], G[2], rs2, xo2;
    {
        const int cg = C / 8, g = (quad * 4) / cg;
        const float mean = a.stats[((size_t)b * 8 + g) * 2 + 0];
        const float rstd = a.stats[((size_t)b * 8 + g) * 2 + 1];
        rs2 = pkf2(rstd, rstd);
        xo2 = pkf2(-mean * rstd, -mean * rstd);
        const float* fr = a.film + (size_t)b * a.film_b_stride + a.film_off;
        float a_[4], b_[4], g_[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = quad * 4 + i;
            const float aa = rstd * a.gn_w[c];
            a_[i] = 0.5f * aa;
            b_[i] = 0.5f * (a.gn_b[c] - mean * aa);
            g_[i] = 1.0f + fr[c];
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            hA[h] = pkf2(a_[2 * h], a_[2 * h + 1]);
            hB[h] = pkf2(b_[2 * h], b_[2 * h + 1]);
            G[h] = pkf2(g_[2 * h], g_[2 * h + 1]);
        }
    }
    f32x2 acc[2][NV];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[h][v] = 0ull;
    const bf16* raw = (const bf16*)a.raw + (size_t)b * L * C + quad * 4;
    const bf16* doa = a.do_a ? (const bf16*)a.do_a + (size_t)b * L * C + quad * 4 : nullptr;
    const int Lp = L / 2;
    const bf16* dop = a.do_pool ? (const bf16*)a.do_pool + (size_t)b * Lp * C + quad * 4 : nullptr;
    const float* cbase = a.cond + (size_t)b * L * Cc;
    const int r0 = blockIdx.x * a.rows_per_cta;
    const int r_end = min(r0 + a.rows_per_cta, L);
    const f32x2 half2 = pkf2(0.5f, 0.5f);
    const float* de = (HEAD && a.do_eps) ? a.do_eps + (size_t)b * L : nullptr;
    f32x2 wk[3][2] = {{0ull, 0ull}, {0ull, 0ull}, {0ull, 0ull}};
    if (HEAD && de) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                wk[k][h] = pkf2(a.do_w[(quad * 4 + 2 * h) * 3 + k], a.do_w[(quad * 4 + 2 * h + 1) * 3 + k]);
    }
    constexpr int UN = HEAD ? 2 : 4;
    for (int r = r0 + tr; r < r_end; r += n_tr * UN) {
        f32x2 x[UN][2], d[UN][2], pl[UN][2];
        float cv[UN][NCA], ev[UN][3];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int rr = r + u * n_tr;
            const int rc = rr < r_end ? rr : r;
            bf16x4_to_pairs(raw + (size_t)rc * C, x[u]);
            if (doa) bf16x4_to_pairs(doa + (size_t)rc * C, d[u]);
            if (dop) bf16x4_to_pairs(dop + (size_t)min(rc >> 1, Lp - 1) * C, pl[u]);
            if (HEAD && de) {
                ev[u][0] = rc + 1 < L ? de[rc + 1] : 0.0f;
                ev[u][1] = de[rc];
                ev[u][2] = rc > 0 ? de[rc - 1] : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < NCA; ++j) cv[u][j] = (NC > 0 && j < Cc) ? cbase[(size_t)rc * Cc + j] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int rr = r + u * n_tr;
            if (rr >= r_end) break;
            const bool pool_ok = dop != nullptr && (rr >> 1) < Lp;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                f32x2 dv = doa ? d[u][h] : 0ull;
                if (pool_ok) dv = ffma2(pl[u][h], half2, dv);
                if (HEAD && de) {
                    const f32x2 t = ffma2(wk[1][h], pkf2(ev[u][1], ev[u][1]), fmul2(wk[2][h], pkf2(ev[u][2], ev[u][2])));
                    dv = fadd2(dv, ffma2(wk[0][h], pkf2(ev[u][0], ev[u][0]), t));
                }
                f32x2 z, act, dact;
                silu_pair(x[u][h], hA[h], hB[h], z, act, dact);
                const f32x2 dn = fmul2(fmul2(dv, G[h]), dact);
                const f32x2 xh = ffma2(x[u][h], rs2, xo2);
                acc[h][0] = fadd2(acc[h][0], dv);
                acc[h][1] = ffma2(dv, act, acc[h][1]);
                acc[h][2] = fadd2(acc[h][2], dn);
                acc[h][3] = ffma2(dn, xh, acc[h][3]);
#pragma unroll
                for (int j = 0; j < NC; ++j) acc[h][4 + j] = ffma2(dv, pkf2(cv[u][j], cv[u][j]), acc[h][4 + j]);
            }
        }
    }
    const int nvr = 4 + Cc;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int v = 0; v < NV; ++v)
            if (v < nvr) {
                float lo, hi;
                upk2(acc[h][v], lo, hi);
                red[((size_t)tr * C + quad * 4 + 2 * h) * nvr + v] = lo;
                red[((size_t)tr * C + quad * 4 + 2 * h + 1) * nvr + v] = hi;
            }
    __syncthreads();
    float* pt = partial + ((size_t)b * gridDim.x + blockIdx.x) * C * nvr;
    for (int i = threadIdx.x; i < C * nvr; i += 256) {
        float sacc = 0.0f;
        for (int t = 0; t < n_tr; ++t) sacc += red[(size_t)t * C * nvr + i];
        pt[i] = sacc;
    }
}

template <bool HEAD>
__global__ void __launch_bounds__(256, 2) gn_bwd_apply_bf16_kernel(GnBwdArgs a, const float* __restrict__ gstat,
                                                                   bf16* __restrict__ d_raw, float* __restrict__ partial_bias) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ float red[];                      // [n_tr][C]
    const int b = blockIdx.y, C = a.C, L = a.L;
    const int n_quad = C / 4, n_tr = 256 / n_quad;
    const int quad = threadIdx.x % n_quad, tr = threadIdx.x / n_quad;
    f32x2 hA[2], hB[2], G[2], GW2[2], rs2, xo2, nm1, nm2;
    {
        const int cg = C / 8, g = (quad * 4) / cg;
        const float mean = a.stats[((size_t)b * 8 + g) * 2 + 0];
        const float rstd = a.stats[((size_t)b * 8 + g) * 2 + 1];
        const float m1 = gstat[((size_t)b * 8 + g) * 2 + 0], m2 = gstat[((size_t)b * 8 + g) * 2 + 1];
        rs2 = pkf2(rstd, rstd);
        xo2 = pkf2(-mean * rstd, -mean * rstd);
        nm1 = pkf2(-m1, -m1);
        nm2 = pkf2(-m2, -m2);
        const float* fr = a.film + (size_t)b * a.film_b_stride + a.film_off;
        float a_[4], b_[4], g_[4], w_[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = quad * 4 + i;
            w_[i] = a.gn_w[c];
            const float aa = rstd * w_[i];
            a_[i] = 0.5f * aa;
            b_[i] = 0.5f * (a.gn_b[c] - mean * aa);
            g_[i] = 1.0f + fr[c];
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            hA[h] = pkf2(a_[2 * h], a_[2 * h + 1]);
            hB[h] = pkf2(b_[2 * h], b_[2 * h + 1]);
            G[h] = pkf2(g_[2 * h], g_[2 * h + 1]);
            GW2[h] = pkf2(w_[2 * h], w_[2 * h + 1]);
        }
    }
    f32x2 sb[2] = {0ull, 0ull};
    const bf16* raw = (const bf16*)a.raw + (size_t)b * L * C + quad * 4;
    const bf16* doa = a.do_a ? (const bf16*)a.do_a + (size_t)b * L * C + quad * 4 : nullptr;
    const int Lp = L / 2;
    const bf16* dop = a.do_pool ? (const bf16*)a.do_pool + (size_t)b * Lp * C + quad * 4 : nullptr;
    bf16* out = d_raw + (size_t)b * L * C + quad * 4;
    const int r0 = blockIdx.x * a.rows_per_cta;
    const int r_end = min(r0 + a.rows_per_cta, L);
    const f32x2 half2 = pkf2(0.5f, 0.5f);
    const float* de = (HEAD && a.do_eps) ? a.do_eps + (size_t)b * L : nullptr;
    f32x2 wk[3][2] = {{0ull, 0ull}, {0ull, 0ull}, {0ull, 0ull}};
    if (HEAD && de) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                wk[k][h] = pkf2(a.do_w[(quad * 4 + 2 * h) * 3 + k], a.do_w[(quad * 4 + 2 * h + 1) * 3 + k]);
    }
    constexpr int UN = HEAD ? 2 : 4;
    for (int r = r0 + tr; r < r_end; r += n_tr * UN) {
        f32x2 x[UN][2], d[UN][2], pl[UN][2];
        float ev[UN][3];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int rr = r + u * n_tr;
            const int rc = rr < r_end ? rr : r;
            bf16x4_to_pairs(raw + (size_t)rc * C, x[u]);
            if (doa) bf16x4_to_pairs(doa + (size_t)rc * C, d[u]);
            if (dop) bf16x4_to_pairs(dop + (size_t)min(rc >> 1, Lp - 1) * C, pl[u]);
            if (HEAD && de) {
                ev[u][0] = rc + 1 < L ? de[rc + 1] : 0.0f;
                ev[u][1] = de[rc];
                ev[u][2] = rc > 0 ? de[rc - 1] : 0.0f;
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int rr = r + u * n_tr;
            if (rr >= r_end) break;
            const bool pool_ok = dop != nullptr && (rr >> 1) < Lp;
            uint2 st;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                f32x2 dv = doa ? d[u][h] : 0ull;
                if (pool_ok) dv = ffma2(pl[u][h], half2, dv);
                if (HEAD && de) {
                    const f32x2 t = ffma2(wk[1][h], pkf2(ev[u][1], ev[u][1]), fmul2(wk[2][h], pkf2(ev[u][2], ev[u][2])));
                    dv = fadd2(dv, ffma2(wk[0][h], pkf2(ev[u][0], ev[u][0]), t));
                }
                f32x2 z, act, dact;
                silu_pair(x[u][h], hA[h], hB[h], z, act, dact);
                const f32x2 dn = fmul2(fmul2(dv, G[h]), dact);
                const f32x2 xh = ffma2(x[u][h], rs2, xo2);
                const f32x2 dz = fmul2(ffma2(xh, nm2, ffma2(dn, GW2[h], nm1)), rs2);
                sb[h] = fadd2(sb[h], dz);
                float lo, hi;
                upk2(dz, lo, hi);
                if (h == 0) st.x = pack_bf16x2(lo, hi); else st.y = pack_bf16x2(lo, hi);
            }
            *reinterpret_cast<uint2*>(out + (size_t)rr * C) = st;
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float lo, hi;
        upk2(sb[h], lo, hi);
        red[(size_t)tr * C + quad * 4 + 2 * h] = lo;
        red[(size_t)tr * C + quad * 4 + 2 * h + 1] = hi;
    }
    __syncthreads();
    float* pt = partial_bias + ((size_t)b * gridDim.x + blockIdx.x) * C;
    for (int i = threadIdx.x; i < C; i += 256) {
        float sacc = 0.0f;
        for (int t = 0; t < n_tr; ++t) sacc += red[(size_t)t * C + i];
        pt[i] = sacc;
    }
}

int g_gn_bwd_stream = 1;
extern int g_final_stream;
extern int g_pdl, g_conv_in_mma, g_wgrad_in_mma, g_cgn_pair2, g_cgn_one_group;
extern int g_gn_bwd_fused, g_gn_bwd_fused_slice;
extern int g_gn_bwd_stats_fast;
extern "C" int gw_set_option(const char* name, int value) {
    if (strcmp(name, "gn_bwd_stream") == 0) { g_gn_bwd_stream = value; return GW_OK; }
    if (strcmp(name, "gn_bwd_stats_fast") == 0) { g_gn_bwd_stats_fast = value; return GW_OK; }
    if (strcmp(name, "gn_bwd_fused") == 0) { g_gn_bwd_fused = value; return GW_OK; }
    if (strcmp(name, "gn_bwd_fused_slice") == 0) { g_gn_bwd_fused_slice = value; return GW_OK; }
    if (strcmp(name, "final_stream") == 0) { g_final_stream = value; return GW_OK; }
    if (strcmp(name, "pdl") == 0) { g_pdl = value; return GW_OK; }
    if (strcmp(name, "one_group") == 0) { g_cgn_one_group = value; return GW_OK; }
    if (strcmp(name, "pair2") == 0) { g_cgn_pair2 = value; return GW_OK; }
    if (strcmp(name, "final_bwd_stream") == 0) { g_final_bwd_stream = value; return GW_OK; }
    if (strcmp(name, "conv_in_mma") == 0) { g_conv_in_mma = value; return GW_OK; }
    if (strcmp(name, "wgrad_in_mma") == 0) { g_wgrad_in_mma = value; return GW_OK; }
    gw_set_error("gw_set_option: unknown option %s", name);
    return GW_ERR_ARG;
}

static int gn_rows_per_cta(int L, int C) {
    const int n_tr = 256 / (C / 4);
    int rows = 32 * n_tr;
    if (rows > L) rows = L;
    return rows < 1 ? 1 : rows;
}

extern "C" long gw_gn_bwd_scratch_elems(int B, int L, int C, int Cc) {
    const int rows = gn_rows_per_cta(L, C), n_rc = gw_cdiv(L, rows), nvr = 5 + Cc;     // (+1: sum of xhat, fast sums kernel)
    // [partials | per-sample reduced | group means]
    const long two_pass = (long)B * n_rc * C * nvr + (long)B * C * nvr + (long)B * 16;
    // one-pass kernel (gn_bwd_fused.cu): up to XCHG_MAX_G row CTAs per sample, plus the conv-bias partials
    const long one_pass = (long)B * XCHG_MAX_G * C * nvr + (long)B * C * nvr + (long)B * 16 + (long)B * XCHG_MAX_G * C;
    return two_pass > one_pass ? two_pass : one_pass;
}
extern "C" long gw_gn_bwd_sync_bytes(int B) { return 64 + (long)B * XCHG_MAX_G * 16 * 8; }
// CTAs per sample the one-pass kernel would use for this layer (0: the two-pass kernels run)
extern "C" int gw_gn_bwd_fused_group(int L, int C, int Cc, int has_do, int has_pool) {
    return gn_bwd_fused_group(L, C, Cc, has_do != 0, has_pool != 0);
}

template <typename T, bool FAST>
static int gn_bwd_run(const GnBwdArgs& a, int B, float* scratch, float* dfilm, long dfilm_b_stride, void* d_raw, float* d_gn_w,
                      float* d_gn_b, float* d_wc, float* d_bc, float* d_conv_bias, void* sync, cudaStream_t st, int phase = 0) {
    // phase (gw_gn_bwd_phase): 0 = everything; 1 = everything but the parameter-gradient kernel where that kernel is the LAST launch
    // (the specialised streaming path), 2 = only that kernel (a no-op on the other paths, where phase 1 already ran it): lets the
    // caller take the small parameter reduction off the critical path (a second stream) -- d_raw is complete after phase 1
    const int C = a.C, L = a.L, Cc = a.Cc, nvr = 4 + Cc;
    if (phase != 0) sync = nullptr;
    if (FAST && sync != nullptr && a.do_eps == nullptr) {
        // one-pass kernel: operands read once, d_raw formed out of shared memory (gn_bwd_fused.cu)
        const int G = gn_bwd_fused_group(L, C, Cc, a.do_a != nullptr, a.do_pool != nullptr);
        if (G > 0) {
            float* partial = scratch;
            float* redb = partial + (size_t)B * G * C * nvr;
            float* gstat = redb + (size_t)B * C * nvr;
            float* biasp = gstat + (size_t)B * 16;
            const int rcf = gn_bwd_fused(a, B, partial, biasp, d_raw, sync, st);
            if (rcf == GW_OK) {
                GW_CUDA(gw_launch_pdl(gn_bwd_finalize_kernel, dim3(B), dim3(1024), (size_t)((size_t)C * nvr * sizeof(float)), st, partial, G, C, nvr, Cc, L, a.gn_w, a.wc, a.bc,
                                                                                        redb, dfilm, dfilm_b_stride, a.film_off, gstat));
                GW_LAUNCH_CHECK();
                GW_CUDA(gw_launch_pdl(gn_bwd_param_kernel, dim3(gw_cdiv(C, 32)), dim3(1024), (size_t)(0), st, redb, B, C, nvr, Cc, a.film, a.film_b_stride, a.film_off, d_gn_w,
                                                                   d_gn_b, d_wc, d_bc, nullptr, 0, nullptr));
                GW_LAUNCH_CHECK();
                return reduce_rows(biasp, B * G, C, C, 1.0f, d_conv_bias, 1, st);
            }
            if (rcf != GW_ERR_UNSUPPORTED) return rcf;
        }
    }
    const int n_rc = gw_cdiv(L, a.rows_per_cta), n_tr = 256 / (C / 4);
    // bf16 with friendly shapes: HBM-streaming kernels (stream_gn.cu), same partial layouts
    const bool stream_ok = FAST && g_gn_bwd_stream && L % 4 == 0 && (C == 64 || C == 128 || C == 256) &&
                           a.rows_per_cta == gn_bwd_stream_rows(L, C);
    // compile-time-specialised streaming kernels: the parameter kernel runs after the apply pass and folds its conv-bias
    // partials too (no reduce launches)
    const bool sx = stream_ok && gn_bwd_stream_fast_ok(a);
    const int nvs = nvr;
    float* partial = scratch;
    float* redb = partial + (size_t)B * n_rc * C * nvs;
    float* gstat = redb + (size_t)B * C * nvs;
    float* biasp = gstat + (size_t)B * 16;                   // [B * n_rc, C] (sx)
    if (phase == 2) {
        if (!sx) return GW_OK;
        GW_CUDA(gw_launch_pdl(gn_bwd_param_kernel, dim3(gw_cdiv(C, 32)), dim3(1024), (size_t)(0), st, redb, B, C, nvs, Cc, a.film, a.film_b_stride, a.film_off, d_gn_w, d_gn_b,
                              d_wc, d_bc, biasp, n_rc, d_conv_bias));
        GW_LAUNCH_CHECK();
        return GW_OK;
    }
    dim3 grid(n_rc, B);
    const size_t sm1 = (size_t)n_tr * C * nvr * sizeof(float), sm2 = (size_t)n_tr * C * sizeof(float);
#define GNB_GO(CCV)                                                                                                       \
    do {                                                                                                                  \
        if (FAST && stream_ok) {                                                                                          \
            int rcs = gn_bwd_stats_stream(a, B, partial, st);                                                              \
            if (rcs != GW_OK) return rcs;                                                                                 \
        } else if (FAST && a.do_eps != nullptr) {                                                                         \
            GW_CUDA(cudaFuncSetAttribute(gn_bwd_stats_bf16_kernel<CCV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1)); \
            GW_CUDA(gw_launch_pdl(gn_bwd_stats_bf16_kernel<CCV, true>, grid, dim3(256), (size_t)(sm1), st, a, partial));                                      \
        } else if (FAST) {                                                                                                \
            GW_CUDA(cudaFuncSetAttribute(gn_bwd_stats_bf16_kernel<CCV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1)); \
            GW_CUDA(gw_launch_pdl(gn_bwd_stats_bf16_kernel<CCV, false>, grid, dim3(256), (size_t)(sm1), st, a, partial));                                     \
        } else {                                                                                                          \
            GW_CUDA(cudaFuncSetAttribute(gn_bwd_stats_kernel<T, FAST, CCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1)); \
            GW_CUDA(gw_launch_pdl(gn_bwd_stats_kernel<T, FAST, CCV>, grid, dim3(256), (size_t)(sm1), st, a, partial));                                        \
        }                                                                                                                 \
    } while (0)
    if (Cc == 0) GNB_GO(0);
    else if (Cc == 1) GNB_GO(1);
    else if (Cc == 5) GNB_GO(5);
    else GNB_GO(-1);
#undef GNB_GO
    GW_LAUNCH_CHECK();
    GW_CUDA(gw_launch_pdl(gn_bwd_finalize_kernel, dim3(B), dim3(1024), (size_t)((size_t)C * nvs * sizeof(float)), st, partial, n_rc, C, nvs, Cc, L, a.gn_w, a.wc, a.bc, redb,
                                                                            dfilm, dfilm_b_stride, a.film_off, gstat));
    GW_LAUNCH_CHECK();
    if (sx) {
        int rcs = gn_bwd_apply_stream(a, B, gstat, d_raw, biasp, st);
        if (rcs != GW_OK) return rcs;
        if (phase == 1) return GW_OK;
        GW_CUDA(gw_launch_pdl(gn_bwd_param_kernel, dim3(gw_cdiv(C, 32)), dim3(1024), (size_t)(0), st, redb, B, C, nvs, Cc, a.film, a.film_b_stride, a.film_off, d_gn_w, d_gn_b,
                                                           d_wc, d_bc, biasp, n_rc, d_conv_bias));
        GW_LAUNCH_CHECK();
        return GW_OK;
    }
    GW_CUDA(gw_launch_pdl(gn_bwd_param_kernel, dim3(gw_cdiv(C, 32)), dim3(1024), (size_t)(0), st, redb, B, C, nvs, Cc, a.film, a.film_b_stride, a.film_off, d_gn_w, d_gn_b,
                                                       d_wc, d_bc, nullptr, 0, nullptr));
    GW_LAUNCH_CHECK();
    // the apply pass reuses the partial region for the conv-bias partials ([B*n_rc, C] <= [B*n_rc, C*nvr])
#define GNA_GO(CCV) GW_CUDA(gw_launch_pdl(gn_bwd_apply_kernel<T, FAST, CCV>, grid, dim3(256), (size_t)(sm2), st, a, gstat, (T*)d_raw, partial))
    if (FAST && stream_ok) {
        int rcs = gn_bwd_apply_stream(a, B, gstat, d_raw, partial, st);
        if (rcs != GW_OK) return rcs;
    } else if (FAST && a.do_eps != nullptr) GW_CUDA(gw_launch_pdl(gn_bwd_apply_bf16_kernel<true>, grid, dim3(256), (size_t)(sm2), st, a, gstat, (bf16*)d_raw, partial));
    else if (FAST) GW_CUDA(gw_launch_pdl(gn_bwd_apply_bf16_kernel<false>, grid, dim3(256), (size_t)(sm2), st, a, gstat, (bf16*)d_raw, partial));
    else if (Cc == 0) GNA_GO(0);
    else if (Cc == 1) GNA_GO(1);
    else if (Cc == 5) GNA_GO(5);
    else GNA_GO(-1);
#undef GNA_GO
    GW_LAUNCH_CHECK();
    return reduce_rows(partial, B * n_rc, C, C, 1.0f, d_conv_bias, 1, st);
}

// d_raw [B, L, C] (dtype); parameter gradients are ACCUMULATED into d_gn_w, d_gn_b [C], d_wc [C, Cc], d_bc [C],
// d_conv_bias [C]; dfilm row b gets (d gamma | d beta) of this layer at film_off (overwritten).
// gw_gn_bwd2: sync_buf (gw_gn_bwd_sync_bytes(B) bytes, zeroed once by the caller, kept across calls) enables the one-pass
// bf16 kernel where the layer shape allows it; NULL = two-pass kernels.
extern "C" int gw_gn_bwd2(const void* raw, const float* stats, int B, int L, int C, const float* gn_w, const float* gn_b,
                          const float* cond, int Cc, const float* wc, const float* bc, const float* film, int film_off,
                          long film_b_stride, const void* do_a, const void* do_pool, const float* do_eps, const float* do_w,
                          int dtype, float* scratch, float* dfilm, long dfilm_b_stride, void* d_raw, float* d_gn_w,
                          float* d_gn_b, float* d_wc, float* d_bc, float* d_conv_bias, void* sync_buf, void* stream) {
    GW_REQUIRE(C % 64 == 0 && C <= 1024 && 256 % (C / 4) == 0, "gw_gn_bwd: C=%d", C);
    GW_REQUIRE(Cc >= 0 && Cc <= BW_MAX_CC, "gw_gn_bwd: Cc=%d", Cc);
    GW_REQUIRE((cond != nullptr) == (Cc > 0), "gw_gn_bwd: cond/Cc mismatch");
    GW_REQUIRE(do_a != nullptr || do_pool != nullptr || do_eps != nullptr, "gw_gn_bwd: no incoming gradient");
    GW_REQUIRE(do_eps == nullptr || (dtype == GW_BF16 && do_w != nullptr), "gw_gn_bwd: the head-gradient source needs bf16 and do_w");
    GW_REQUIRE(dtype == GW_F32 || dtype == GW_BF16, "gw_gn_bwd: dtype %d", dtype);
    GnBwdArgs a;
    a.raw = raw; a.stats = stats; a.gn_w = gn_w; a.gn_b = gn_b; a.cond = cond; a.wc = wc; a.bc = bc; a.film = film;
    a.film_b_stride = film_b_stride; a.film_off = film_off; a.do_a = do_a; a.do_pool = do_pool; a.do_eps = do_eps; a.do_w = do_w;
    a.L = L; a.C = C; a.Cc = Cc;
    a.rows_per_cta = gn_rows_per_cta(L, C);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GW_F32)
        return gn_bwd_run<float, false>(a, B, scratch, dfilm, dfilm_b_stride, d_raw, d_gn_w, d_gn_b, d_wc, d_bc, d_conv_bias, nullptr, st);
    return gn_bwd_run<bf16, true>(a, B, scratch, dfilm, dfilm_b_stride, d_raw, d_gn_w, d_gn_b, d_wc, d_bc, d_conv_bias, sync_buf, st);
}
extern "C" int gw_gn_bwd_phase(const void* raw, const float* stats, int B, int L, int C, const float* gn_w, const float* gn_b,
                               const float* cond, int Cc, const float* wc, const float* bc, const float* film, int film_off,
                               long film_b_stride, const void* do_a, const void* do_pool, const float* do_eps, const float* do_w,
                               int dtype, float* scratch, float* dfilm, long dfilm_b_stride, void* d_raw, float* d_gn_w,
                               float* d_gn_b, float* d_wc, float* d_bc, float* d_conv_bias, int phase, void* stream) {
    GW_REQUIRE(C % 64 == 0 && C <= 1024 && 256 % (C / 4) == 0, "gw_gn_bwd: C=%d", C);
    GW_REQUIRE(Cc >= 0 && Cc <= BW_MAX_CC, "gw_gn_bwd: Cc=%d", Cc);
    GW_REQUIRE((cond != nullptr) == (Cc > 0), "gw_gn_bwd: cond/Cc mismatch");
    GW_REQUIRE(do_a != nullptr || do_pool != nullptr || do_eps != nullptr, "gw_gn_bwd: no incoming gradient");
    GW_REQUIRE(do_eps == nullptr || (dtype == GW_BF16 && do_w != nullptr), "gw_gn_bwd: the head-gradient source needs bf16 and do_w");
    GW_REQUIRE(dtype == GW_F32 || dtype == GW_BF16, "gw_gn_bwd: dtype %d", dtype);
    GW_REQUIRE(phase >= 0 && phase <= 2, "gw_gn_bwd_phase: phase %d", phase);
    GnBwdArgs a;
    a.raw = raw; a.stats = stats; a.gn_w = gn_w; a.gn_b = gn_b; a.cond = cond; a.wc = wc; a.bc = bc; a.film = film;
    a.film_b_stride = film_b_stride; a.film_off = film_off; a.do_a = do_a; a.do_pool = do_pool; a.do_eps = do_eps; a.do_w = do_w;
    a.L = L; a.C = C; a.Cc = Cc;
    a.rows_per_cta = gn_rows_per_cta(L, C);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GW_F32)
        return gn_bwd_run<float, false>(a, B, scratch, dfilm, dfilm_b_stride, d_raw, d_gn_w, d_gn_b, d_wc, d_bc, d_conv_bias, nullptr, st, phase);
    return gn_bwd_run<bf16, true>(a, B, scratch, dfilm, dfilm_b_stride, d_raw, d_gn_w, d_gn_b, d_wc, d_bc, d_conv_bias, nullptr, st, phase);
}
extern "C" int gw_gn_bwd(const void* raw, const float* stats, int B, int L, int C, const float* gn_w, const float* gn_b,
                         const float* cond, int Cc, const float* wc, const float* bc, const float* film, int film_off,
                         long film_b_stride, const void* do_a, const void* do_pool, const float* do_eps, const float* do_w,
                         int dtype, float* scratch, float* dfilm, long dfilm_b_stride, void* d_raw, float* d_gn_w,
                         float* d_gn_b, float* d_wc, float* d_bc, float* d_conv_bias, void* stream) {
    return gw_gn_bwd2(raw, stats, B, L, C, gn_w, gn_b, cond, Cc, wc, bc, film, film_off, film_b_stride, do_a, do_pool, do_eps, do_w,
                      dtype, scratch, dfilm, dfilm_b_stride, d_raw, d_gn_w, d_gn_b, d_wc, d_bc, d_conv_bias, nullptr, stream);
}

// ------------------------------------------------------------------------------------------------
// exact-mode conv backward helpers
// ------------------------------------------------------------------------------------------------
// dgrad of Conv1d(k=3, pad=1) is the same conv with w'[ci][co][k] = w[co][ci][2-k]  (run through gw_conv3_simt)
__global__ void __launch_bounds__(256) weight_dgrad_kernel(const float* __restrict__ w, int Cout, int Cin, float* __restrict__ wt) {
    pdl_wait();
    pdl_launch_dependents();
    const long n = (long)Cout * Cin * 3;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i % 3);
        const long r = i / 3;
        const int co = (int)(r % Cout), ci = (int)(r / Cout);
        wt[i] = w[((size_t)co * Cin + ci) * 3 + (2 - k)];
    }
}
extern "C" int gw_weight_dgrad(const float* w, int Cout, int Cin, float* wt, void* stream) {
    GW_REQUIRE(Cout > 0 && Cin > 0, "gw_weight_dgrad: sizes");
    const long n = (long)Cout * Cin * 3;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    GW_CUDA(gw_launch_pdl(weight_dgrad_kernel, grid, dim3(256), (size_t)(0), (cudaStream_t)stream, w, Cout, Cin, wt));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// gradient of cat[nearest-upsample x2 (h), skip] (models.py:217-221): d_h[m] = d_cat[2m, :C0] + d_cat[2m+1, :C0],
// d_skip = d_cat[:, C0:].  d_cat [B, L, C0+C1]; d_h [B, L0, C0]; d_skip [B, L, C1].
template <typename T>
__global__ void __launch_bounds__(256) split_cat_grad_kernel(const T* __restrict__ d_cat, int L, int C0, int L0, int C1,
                                                             T* __restrict__ d_h, T* __restrict__ d_skip) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.y;
    const int Ct = C0 + C1;
    const int q0 = C0 / 4, q1 = C1 / 4;
    const long n_h = (long)L0 * q0, n_s = (long)L * q1;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_h + n_s; i += (long)gridDim.x * blockDim.x) {
        if (i < n_h) {
            const int m = (int)(i / q0), q = (int)(i % q0);
            float a[4] = {0.0f, 0.0f, 0.0f, 0.0f}, c[4];
            if (2 * m < L) ld4f(d_cat + ((size_t)b * L + 2 * m) * Ct + q * 4, a);
            if (2 * m + 1 < L) {
                ld4f(d_cat + ((size_t)b * L + 2 * m + 1) * Ct + q * 4, c);
#pragma unroll
                for (int j = 0; j < 4; ++j) a[j] += c[j];
            }
            st4f(d_h + ((size_t)b * L0 + m) * C0 + q * 4, a);
        } else {
            const long k = i - n_h;
            const int l = (int)(k / q1), q = (int)(k % q1);
            float a[4];
            ld4f(d_cat + ((size_t)b * L + l) * Ct + C0 + q * 4, a);
            st4f(d_skip + ((size_t)b * L + l) * C1 + q * 4, a);
        }
    }
}
extern "C" int gw_split_cat_grad(const void* d_cat, int B, int L, int C0, int L0, int C1, void* d_h, void* d_skip, int dtype,
                                 void* stream) {
    GW_REQUIRE(C0 % 4 == 0 && C1 % 4 == 0 && C0 > 0 && C1 > 0, "gw_split_cat_grad: channels");
    GW_REQUIRE(dtype == GW_F32 || dtype == GW_BF16, "gw_split_cat_grad: dtype %d", dtype);
    const long n = (long)L0 * (C0 / 4) + (long)L * (C1 / 4);
    int gx = (int)((n + 255) / 256);
    if (gx > 1024) gx = 1024;
    dim3 grid(gx, B);
    if (dtype == GW_F32)
        GW_CUDA(gw_launch_pdl(split_cat_grad_kernel<float>, grid, dim3(256), (size_t)(0), (cudaStream_t)stream, (const float*)d_cat, L, C0, L0, C1, (float*)d_h, (float*)d_skip));
    else
        GW_CUDA(gw_launch_pdl(split_cat_grad_kernel<bf16>, grid, dim3(256), (size_t)(0), (cudaStream_t)stream, (const bf16*)d_cat, L, C0, L0, C1, (bf16*)d_h, (bf16*)d_skip));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// wgrad of Conv1d(k=3, pad=1) over the virtual concat [nearest-upsample(src0) | src1], CUDA-core fp32:
//   dW[co][ci][k] = sum_{b,l} d_raw[b,l,co] * in[b, l+k-1, ci]
// CTA = 64 couts x 64 cins x 3 taps over a strided share of (sample, 32-position chunk) work items; the split-K
// partials [n_split][Cout][Cin][3] land in scratch and are summed in fixed order into dW.
template <typename T>
__global__ void __launch_bounds__(256) wgrad3_simt_kernel(const T* __restrict__ src0, int C0, int L0, int up0,
                                                          const T* __restrict__ src1, int C1, const T* __restrict__ d_raw, int B,
                                                          int L, int Cout, float* __restrict__ partial) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int KP = 32;
    __shared__ __align__(16) float dy[KP][64];
    __shared__ __align__(16) float xs[KP + 2][64];
    const int Cin = C0 + C1;
    const int ci0 = blockIdx.x * 64, co0 = blockIdx.y * 64, split = blockIdx.z, n_split = gridDim.z;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[i][j][k] = 0.0f;
    const int n_chunk = (L + KP - 1) / KP;
    const long total = (long)B * n_chunk;
    const bool from0 = ci0 < C0;
    for (long w = split; w < total; w += n_split) {
        const int b = (int)(w / n_chunk), l0 = (int)(w % n_chunk) * KP;
        {
            const int row = threadIdx.x >> 3, oct = threadIdx.x & 7;
            const int l = l0 + row;
            float v[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            if (l < L) ld8(d_raw + ((size_t)b * L + l) * Cout + co0 + oct * 8, v);
            *reinterpret_cast<float4*>(&dy[row][oct * 8]) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(&dy[row][oct * 8 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
        }
        for (int i = threadIdx.x; i < (KP + 2) * 8; i += 256) {
            const int row = i >> 3, oct = i & 7;
            const int l = l0 + row - 1;
            float v[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            if (l >= 0 && l < L) {
                if (from0) {
                    const int ls = up0 ? (l >> 1) : l;
                    if (ls < L0) ld8(src0 + ((size_t)b * L0 + ls) * C0 + ci0 + oct * 8, v);
                } else {
                    ld8(src1 + ((size_t)b * L + l) * C1 + (ci0 - C0) + oct * 8, v);
                }
            }
            *reinterpret_cast<float4*>(&xs[row][oct * 8]) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(&xs[row][oct * 8 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
        }
        __syncthreads();
#pragma unroll 4
        for (int p = 0; p < KP; ++p) {
            const float4 d4 = *reinterpret_cast<const float4*>(&dy[p][ty * 4]);
            const float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 x4 = *reinterpret_cast<const float4*>(&xs[p + k][tx * 4]);
                const float x[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j][k] = fmaf(d[i], x[j], acc[i][j][k]);
            }
        }
        __syncthreads();
    }
    float* pt = partial + (size_t)split * Cout * Cin * 3;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                pt[((size_t)(co0 + ty * 4 + i) * Cin + ci0 + tx * 4 + j) * 3 + k] = acc[i][j][k];
}

extern "C" int gw_wgrad3_simt(const void* src0, int C0, int L0, int up0, const void* src1, int C1, const void* d_raw, int B,
                              int L, int Cout, int dtype, float* scratch, long scratch_elems, float* dW, void* stream) {
    GW_REQUIRE(C0 % 64 == 0 && C1 % 64 == 0 && C0 > 0 && Cout % 64 == 0, "gw_wgrad3_simt: channels C0=%d C1=%d Cout=%d", C0, C1, Cout);
    GW_REQUIRE((src1 != nullptr) == (C1 > 0), "gw_wgrad3_simt: src1/C1 mismatch");
    GW_REQUIRE(dtype == GW_F32 || dtype == GW_BF16, "gw_wgrad3_simt: dtype %d", dtype);
    const int Cin = C0 + C1;
    const long per = (long)Cout * Cin * 3;
    const long total = (long)B * gw_cdiv(L, 32);
    const int tiles = (Cin / 64) * (Cout / 64);
    long n_split = (148L * 4 + tiles - 1) / tiles;
    if (n_split > total) n_split = total;
    if (n_split > scratch_elems / per) n_split = scratch_elems / per;
    if (n_split > 65535) n_split = 65535;
    GW_REQUIRE(n_split >= 1, "gw_wgrad3_simt: scratch too small (%ld < %ld)", scratch_elems, per);
    dim3 grid(Cin / 64, Cout / 64, (unsigned)n_split);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GW_F32)
        GW_CUDA(gw_launch_pdl(wgrad3_simt_kernel<float>, grid, dim3(256), (size_t)(0), st, (const float*)src0, C0, L0, up0, (const float*)src1, C1, (const float*)d_raw, B, L, Cout, scratch));
    else
        GW_CUDA(gw_launch_pdl(wgrad3_simt_kernel<bf16>, grid, dim3(256), (size_t)(0), st, (const bf16*)src0, C0, L0, up0, (const bf16*)src1, C1, (const bf16*)d_raw, B, L, Cout, scratch));
    GW_LAUNCH_CHECK();
    return reduce_rows(scratch, (int)n_split, per, per, 1.0f, dW, 1, st);
}

int wgrad_in_stream(const float* x, int B, int Cx, int L, const void* d_raw, float* scratch, long scratch_elems, int* n_rows,
                    cudaStream_t st);
extern int g_gn_bwd_stream;

// wgrad of the first conv (models.py:204): dW[co][ci][k] = sum d_raw[b,l,co] * x[b,ci,l+k-1], x fp32 [B, Cx, L].
// A thread owns one channel PAIR and walks groups of 4 consecutive rows: the 6 input samples a group needs are two
// vector smem loads per input channel, reused by 4 rows x 3 taps x 2 channels of FMAs.
__device__ __forceinline__ void ld2f(const float* p, float (&v)[2]) {
    const float2 a = *reinterpret_cast<const float2*>(p);
    v[0] = a.x; v[1] = a.y;
}
__device__ __forceinline__ void ld2f(const bf16* p, float (&v)[2]) {
    const uint32_t r = *reinterpret_cast<const uint32_t*>(p);
    v[0] = __uint_as_float(r << 16); v[1] = __uint_as_float(r & 0xffff0000u);
}
template <typename T, int CXM>
__global__ void __launch_bounds__(256) wgrad_in_kernel(const float* __restrict__ x, int Cx, int L, const T* __restrict__ d_raw,
                                                       int C, float* __restrict__ partial, int rows_per_cta) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(16) float sm[];
    const int pitch = rows_per_cta + 8;           // multiple of 4: xs[ci][j] = x[r0 - 1 + j], float4-aligned at j % 4 == 0
    float* xs = sm;                               // [Cx][pitch]
    float* red = xs + Cx * pitch;                 // [n_tr][C][Cx*3]
    const int n_cp = C / 2, n_tr = 256 / n_cp;
    const int b = blockIdx.y, r0 = blockIdx.x * rows_per_cta;
    for (int i = threadIdx.x; i < Cx * pitch; i += 256) {
        const int c = i / pitch, p = i % pitch;
        const int l = r0 + p - 1;
        xs[i] = (l >= 0 && l < L) ? x[((size_t)b * Cx + c) * L + l] : 0.0f;
    }
    __syncthreads();
    const int cp = threadIdx.x % n_cp, tr = threadIdx.x / n_cp;
    float acc[2][CXM * 3];
#pragma unroll
    for (int i = 0; i < CXM * 3; ++i) { acc[0][i] = 0.0f; acc[1][i] = 0.0f; }
    const int r_end = min(rows_per_cta, L - r0);
    const T* dp = d_raw + ((size_t)b * L + r0) * C + cp * 2;
    for (int g = tr * 4; g < r_end; g += n_tr * 4) {
        float d[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (g + u < r_end) ld2f(dp + (size_t)(g + u) * C, d[u]);
            else { d[u][0] = 0.0f; d[u][1] = 0.0f; }
        }
#pragma unroll
        for (int ci = 0; ci < CXM; ++ci) {
            if (ci < Cx) {
                const float4 xa = *reinterpret_cast<const float4*>(xs + ci * pitch + g);
                const float2 xb = *reinterpret_cast<const float2*>(xs + ci * pitch + g + 4);
                const float xv[6] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y};
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        acc[0][ci * 3 + k] = fmaf(d[u][0], xv[u + k], acc[0][ci * 3 + k]);
                        acc[1][ci * 3 + k] = fmaf(d[u][1], xv[u + k], acc[1][ci * 3 + k]);
                    }
            }
        }
    }
    const int nv = Cx * 3;
#pragma unroll
    for (int i = 0; i < CXM * 3; ++i)
        if (i < nv) {
            red[((size_t)tr * C + cp * 2) * nv + i] = acc[0][i];
            red[((size_t)tr * C + cp * 2 + 1) * nv + i] = acc[1][i];
        }
    __syncthreads();
    float* pt = partial + ((size_t)b * gridDim.x + blockIdx.x) * C * nv;
    for (int i = threadIdx.x; i < C * nv; i += 256) {
        float sacc = 0.0f;
        for (int t = 0; t < n_tr; ++t) sacc += red[(size_t)t * C * nv + i];
        pt[i] = sacc;
    }
}

extern "C" int gw_wgrad_in(const float* x, int B, int Cx, int L, const void* d_raw, int C, int dtype, float* scratch,
                           long scratch_elems, float* dW, void* stream) {
    GW_REQUIRE(C >= 64 && C <= 512 && 512 % C == 0, "gw_wgrad_in: C=%d", C);
    GW_REQUIRE(Cx >= 1 && Cx <= 16, "gw_wgrad_in: Cx=%d", Cx);
    GW_REQUIRE(dtype == GW_F32 || dtype == GW_BF16, "gw_wgrad_in: dtype %d", dtype);
    if (dtype == GW_BF16 && C == 64 && L % 4 == 0 && g_gn_bwd_stream) {       // HBM-streaming kernel (stream_gn.cu)
        int n_rows = 0;
        int rcs = wgrad_in_stream(x, B, Cx, L, d_raw, scratch, scratch_elems, &n_rows, (cudaStream_t)stream);
        if (rcs != GW_OK) return rcs;
        return reduce_rows(scratch, n_rows, (long)C * Cx * 3, (long)C * Cx * 3, 1.0f, dW, 1, (cudaStream_t)stream);
    }
    int rows = 1024;
    if (rows > L) rows = (L + 3) & ~3;
    const int n_rc = gw_cdiv(L, rows), nv = Cx * 3;
    GW_REQUIRE((long)B * n_rc * C * nv <= scratch_elems, "gw_wgrad_in: scratch too small");
    const int n_tr = 256 / (C / 2);
    const size_t smem = ((size_t)Cx * (rows + 8) + (size_t)n_tr * C * nv) * sizeof(float);
    dim3 grid(n_rc, B);
    cudaStream_t st = (cudaStream_t)stream;
#define WIN_GO(TT, CXM)                                                                                               \
    do {                                                                                                              \
        GW_CUDA(cudaFuncSetAttribute(wgrad_in_kernel<TT, CXM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        GW_CUDA(gw_launch_pdl(wgrad_in_kernel<TT, CXM>, grid, dim3(256), (size_t)(smem), st, x, Cx, L, (const TT*)d_raw, C, scratch, rows));                \
    } while (0)
    if (dtype == GW_F32) {
        if (Cx <= 4) WIN_GO(float, 4); else if (Cx <= 8) WIN_GO(float, 8); else WIN_GO(float, 16);
    } else {
        if (Cx <= 4) WIN_GO(bf16, 4); else if (Cx <= 8) WIN_GO(bf16, 8); else WIN_GO(bf16, 16);
    }
#undef WIN_GO
    GW_LAUNCH_CHECK();
    return reduce_rows(scratch, B * n_rc, (long)C * nv, (long)C * nv, 1.0f, dW, 1, st);
}
