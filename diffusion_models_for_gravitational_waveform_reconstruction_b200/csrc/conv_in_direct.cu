// FIRST block for inference in ONE pass with ANALYTIC GroupNorm statistics (sm_100a):
//   Conv1d(C_in -> 64, k=3) -> GroupNorm(8) -> SiLU -> + cond 1x1 conv -> FiLM -> out [B, L, 64] bf16 and avg_pool1d(out, 2)
// (models.py:160-173, 188-193, 204-208).
//
// conv_in_gn.cu needs the statistics of a whole sample before it can write one output element, so a group of CTAs convolves
// into shared memory, exchanges sums, and applies out of shared memory (~64 warp instructions per row).  But the first conv
// has K = 3*C_in <= 24 inputs per output: its output moments are a QUADRATIC FORM of the input's lag-(0,1,2) cross products,
//     sum_l raw[c,l]   = L b_c + w_c . X1                      X1[(i,k)]       = sum_l xp_i[l+k-1]
//     sum_l raw[c,l]^2 = L b_c^2 + 2 b_c w_c . X1 + w_c' R w_c  R[(i,k),(j,m)] = sum_l xp_i[l+k-1] xp_j[l+m-1]
// (xp = zero-padded input).  gw_conv_in_direct therefore runs
//   1. in_moments_kernel: one CTA per sample reads the (tiny) fp32 input once, accumulates the C_in^2 x 3 lagged products,
//      folds them with the weights into mean / rstd of the 8 groups and writes the sample's per-channel epilogue coefficients;
//   2. conv_in_direct_kernel: any CTA takes any 256-row slice: tf32 mma.sync conv with fp32 accumulators in registers, the
//      GroupNorm / SiLU / cond / FiLM epilogue straight on the accumulator fragments, stmatrix into a swizzled 16-row tile per
//      warp, 16-byte coalesced stores of out and of the 2:1 pooled rows.  No exchange between CTAs, no co-residency rule, no
//      second pass over shared memory.
// Statistics are those of the exact fp32 conv output (the oracle's), not of a rounded copy.
#include "common.cuh"
#include "../../include/gwb200.h"

#define CID_C 64
#define CID_ROWS 256
#define CID_XP (CID_ROWS + 8)
#define CID_MAX_CX 8
#define CID_MAX_CC 8
#define CID_PF (CID_MAX_CX + 1)         // prefetch registers per thread: one column of every channel + the window's tail
#define CID_CF(nca) ((8 + 2 * (nca) + 3) / 4 * 4)      // floats per channel pair (16-byte aligned rows: read with LDS.128)

struct CidArgs {
    const float* xa;
    const float* xb;
    const int* step_ptr;
    const float* w;
    const float* bias;
    const float* gn_w;
    const float* gn_b;
    const float* wc;
    const float* bc;
    const float* film;
    float* coef;            // [B][32 channel pairs][CID_CF(NCA)]: A0 A1 B0 B1 G0 G1 E0 E1 (W_j0 W_j1)... padded to 16 bytes
    bf16* out;
    bf16* pooled;
    long film_b_stride, film_step_stride;
    int film_off, B, Cx, L, Cc, n_items, slices;
};

__device__ __forceinline__ uint32_t cid_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void cid_mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cid_stmatrix_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------ 1. moments -> coefficients
// One CTA (256 threads) per sample.  The sample's channels are staged chunk by chunk with 1-D bulk copies (no per-element index
// arithmetic); warp w works on channel i = w % Cx and position slab w / Cx of the chunk and accumulates
// q[j][d] = sum_u x_i[u] x_j[u+d], d = 0..2, and T_i = sum_u x_i[u] (fp32 inside a chunk, fp64 across chunks and slabs).
#define CID_CH 2048
__device__ __forceinline__ void cid_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void cid_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cid_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void cid_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
template <int NCA>
__global__ void __launch_bounds__(256) in_moments_kernel(const CidArgs A) {
    constexpr int XPM = CID_CH + 8;
    extern __shared__ __align__(16) unsigned char smem_m[];
    float* xs = reinterpret_cast<float*>(smem_m);                        // [Cx][XPM]
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ double s_q[CID_MAX_CX][CID_MAX_CX][3];
    __shared__ double s_t[CID_MAX_CX];
    __shared__ double s_R[3 * CID_MAX_CX][3 * CID_MAX_CX + 1];
    __shared__ double s_x1[3 * CID_MAX_CX];
    __shared__ float s_edge[CID_MAX_CX][4];                              // x[0], x[1], x[L-2], x[L-1]
    __shared__ double s_s1[CID_C], s_s2[CID_C];
    __shared__ float s_mr[8][2];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Cx = A.Cx, L = A.L, Cc = A.Cc, K = 3 * A.Cx;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    for (int i = tid; i < CID_MAX_CX * CID_MAX_CX * 3; i += 256) (&s_q[0][0][0])[i] = 0.0;
    if (tid < CID_MAX_CX) s_t[tid] = 0.0;
    if (tid == 0) {
        cid_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();
    pdl_launch_dependents();
    const int step = A.step_ptr != nullptr ? *A.step_ptr : 0;
    const float* x = ((step & 1) ? A.xb : A.xa) + (size_t)b * Cx * L;
    const bool bulk_ok = (L % 4) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const int S = 8 / Cx;                                 // position slabs per channel (Cx <= 8)
    const int ci = warp % Cx, slab = warp / Cx;
    const bool active = slab < S;
    uint32_t phase = 0;
    for (int c0 = 0; c0 < L; c0 += CID_CH) {
        const int n = min(CID_CH, L - c0);
        const int ncp = min(n + 4, L - c0);              // staged floats per channel (2 halo positions, rounded to 16 bytes)
        __syncthreads();                                  // the previous chunk's readers are done
        if (bulk_ok) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                cid_mbar_expect_tx(bar, (uint32_t)(Cx * ncp * 4));
                for (int c = 0; c < Cx; ++c)
                    cid_bulk_load((uint32_t)__cvta_generic_to_shared(xs + c * XPM), x + (size_t)c * L + c0, (uint32_t)(ncp * 4), bar);
            }
            if (ncp < n + 2)                              // last chunk: the two halo positions lie beyond the sample
                for (int i = tid; i < Cx * 4; i += 256) xs[(i >> 2) * XPM + ncp + (i & 3)] = 0.0f;
            cid_mbar_wait(bar, phase);
            phase ^= 1;
        } else {
            for (int c = 0; c < Cx; ++c)
                for (int p = tid; p < n + 2; p += 256) xs[c * XPM + p] = c0 + p < L ? x[(size_t)c * L + c0 + p] : 0.0f;
        }
        __syncthreads();
        if (active) {
            float q[CID_MAX_CX][3], t = 0.0f;
#pragma unroll
            for (int j = 0; j < CID_MAX_CX; ++j) q[j][0] = q[j][1] = q[j][2] = 0.0f;
            const float* xi = xs + ci * XPM;
            const int per = (n + S - 1) / S, u0 = slab * per, u1 = min(n, u0 + per);
            for (int u = u0 + lane; u < u1; u += 32) {
                const float v = xi[u];
                t += v;
#pragma unroll
                for (int j = 0; j < CID_MAX_CX; ++j) {
                    if (j < Cx) {
                        const float* xj = xs + j * XPM + u;          // positions beyond the sample are staged as zeros
                        q[j][0] = fmaf(v, xj[0], q[j][0]);
                        q[j][1] = fmaf(v, xj[1], q[j][1]);
                        q[j][2] = fmaf(v, xj[2], q[j][2]);
                    }
                }
            }
            // fold the warp, then add to the CTA totals in fp64 (one atomic per value, per warp, per chunk)
            t = warp_sum(t);
            if (lane == 0) atomicAdd(&s_t[ci], (double)t);
#pragma unroll
            for (int j = 0; j < CID_MAX_CX; ++j) {
                if (j < Cx) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const float v = warp_sum(q[j][d]);
                        if (lane == 0) atomicAdd(&s_q[ci][j][d], (double)v);
                    }
                }
            }
        }
    }
    if (tid < Cx * 4) {
        const int c = tid >> 2, e = tid & 3;
        const int l = e < 2 ? e : L - 4 + e;
        s_edge[c][e] = (l >= 0 && l < L) ? x[(size_t)c * L + l] : 0.0f;
    }
    __syncthreads();
    auto xe = [&](int i, int l) -> double {               // x_i[l] for l in {0, 1, L-2, L-1}
        if (l == 0) return (double)s_edge[i][0];
        if (l == 1) return (double)s_edge[i][1];
        if (l == L - 2) return (double)s_edge[i][2];
        return (double)s_edge[i][3];
    };
    // ---- R[(i,k)][(j,m)] = sum_l xp_i[l+k-1] xp_j[l+m-1] and X1[(i,k)] = sum_l xp_i[l+k-1] from the lagged products:
    //      u = l+k-1 runs over [0, L-2] for k = 0 and [1, L-1] for k = 2, so one boundary product drops out
    for (int e = tid; e < K * K; e += 256) {
        const int a = e / K, c2 = e % K;
        const int i = a / 3, k = a % 3, j = c2 / 3, m = c2 % 3, d = m - k;
        double r = d >= 0 ? s_q[i][j][d] : s_q[j][i][-d];
        if (k == 0 && d <= 0 && L - 1 + d >= 0) r -= xe(i, L - 1) * xe(j, L - 1 + d);
        if (k == 2 && d >= 0 && d < L) r -= xe(i, 0) * xe(j, d);
        s_R[a][c2] = r;
    }
    if (tid < K) {
        const int i = tid / 3, k = tid % 3;
        double x1 = s_t[i];
        if (k == 0) x1 -= xe(i, L - 1);
        if (k == 2) x1 -= xe(i, 0);
        s_x1[tid] = x1;
    }
    __syncthreads();
    // ---- channel c = tid / 4: first and second moment of its conv output; the 4 lanes of a channel split the rows of R
    {
        const int c = tid >> 2, part = tid & 3;
        const float* wr = A.w + (size_t)c * K;
        double lin = 0.0, quad = 0.0;
        for (int a = part; a < K; a += 4) {
            const double wa = (double)wr[a];
            double acc = 0.0;
            for (int c2 = 0; c2 < K; ++c2) acc = fma((double)wr[c2], s_R[a][c2], acc);
            quad = fma(wa, acc, quad);
            lin = fma(wa, s_x1[a], lin);
        }
        lin += __shfl_xor_sync(0xffffffffu, lin, 1);
        quad += __shfl_xor_sync(0xffffffffu, quad, 1);
        lin += __shfl_xor_sync(0xffffffffu, lin, 2);
        quad += __shfl_xor_sync(0xffffffffu, quad, 2);
        if (part == 0) {
            const double bias = (double)A.bias[c];
            s_s1[c] = (double)L * bias + lin;
            s_s2[c] = (double)L * bias * bias + 2.0 * bias * lin + quad;
        }
    }
    __syncthreads();
    if (tid < 8) {
        double a1 = 0.0, a2 = 0.0;
        for (int j = 0; j < 8; ++j) { a1 += s_s1[tid * 8 + j]; a2 += s_s2[tid * 8 + j]; }
        const double inv_n = 1.0 / (8.0 * (double)L);
        const double mean = a1 * inv_n;
        double var = a2 * inv_n - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mr[tid][0] = (float)mean;
        s_mr[tid][1] = (float)(1.0 / sqrt(var + 1e-5));
    }
    __syncthreads();
    // ---- epilogue coefficients of channel c for the conv WITHOUT its bias (acc): h = A acc + B is HALF the GroupNorm output,
    //      out = (h + h tanh h) G + E + sum_j W_j cond_j,  G = 1 + gamma_t, E = bc G + beta_t, W_j = wc_j G
    if (tid < CID_C) {
        const int c = tid, pr = c >> 1, hf = c & 1;
        const float* fr = A.film + (size_t)step * A.film_step_stride + (size_t)b * A.film_b_stride + A.film_off;
        const float mean = s_mr[c >> 3][0], rstd = s_mr[c >> 3][1];
        const float a = rstd * A.gn_w[c];
        const float g = 1.0f + fr[c];
        float* cf = A.coef + ((size_t)b * 32 + pr) * CID_CF(NCA);
        cf[0 + hf] = 0.5f * a;
        cf[2 + hf] = 0.5f * fmaf(A.bias[c] - mean, a, A.gn_b[c]);
        cf[4 + hf] = g;
        cf[6 + hf] = fmaf(Cc > 0 ? A.bc[c] : 0.0f, g, fr[CID_C + c]);
#pragma unroll
        for (int jc = 0; jc < NCA; ++jc) cf[8 + 2 * jc + hf] = jc < Cc ? A.wc[c * Cc + jc] * g : 0.0f;
    }
}

// ------------------------------------------------------------------------------------------------ 2. conv + epilogue
template <int CC>
__global__ void __launch_bounds__(256, 2) conv_in_direct_kernel(const CidArgs A) {
    constexpr int NCA = CC >= 0 ? CC : CID_MAX_CC;       // CC = -1: any count up to 8, zero-padded weights
    constexpr int NCR = NCA > 0 ? NCA : 1;               // register array extent
    constexpr int CF = CID_CF(NCA);                      // floats per channel pair
    constexpr int C = CID_C;
    extern __shared__ __align__(16) unsigned char smem[];
    float* xs = reinterpret_cast<float*>(smem);                           // [Cx + 1][XP]: xs[c][j] = x[c][l00 - 1 + j]; row Cx = zeros
    float* ws = xs + (A.Cx + 1) * CID_XP;                                 // tf32 B fragments [KS][8 n-tiles][32 lanes][2]
    float* cfs = ws + 3 * 8 * 32 * 2;                                     // [32 pairs][CF]
    unsigned char* stg_base = reinterpret_cast<unsigned char*>(cfs + 32 * CF);   // [8 warps][2 KB out tile | 1 KB pooled tile]
    const int Cx = A.Cx, L = A.L, Cc = A.Cc;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int KS = (3 * Cx + 7) / 8;
    // B fragment of (k-step ks, n-tile nt) for lane (g, t4): W[kk = 8 ks + t4 (+4)][co = 8 nt + g]
    for (int i = tid; i < KS * 8 * 32 * 2; i += 256) {
        const int j = i & 1, ln = (i >> 1) & 31, nt = (i >> 6) & 7, ks = i >> 9;
        const int kk = ks * 8 + (ln & 3) + 4 * j, co = nt * 8 + (ln >> 2);
        const float v = kk < 3 * Cx ? A.w[(size_t)co * Cx * 3 + kk] : 0.0f;
        ws[i] = __uint_as_float(cid_tf32(v));
    }
    int aoff[3][2];
#pragma unroll
    for (int ks = 0; ks < 3; ++ks)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int kk = ks * 8 + t4 + 4 * j;
            aoff[ks][j] = kk < 3 * Cx ? (kk / 3) * CID_XP + kk % 3 : Cx * CID_XP;       // padding columns read the zero row
        }
    for (int i = tid; i < CID_XP; i += 256) xs[Cx * CID_XP + i] = 0.0f;
    const uint32_t stg = (uint32_t)__cvta_generic_to_shared(stg_base) + (uint32_t)warp * 3072u;
    const unsigned char* stg_g = stg_base + warp * 3072;
    pdl_wait();                          // the input, the step counter and the coefficient table belong to earlier kernels
    pdl_launch_dependents();
    const int step = A.step_ptr != nullptr ? *A.step_ptr : 0;
    const float* x = (step & 1) ? A.xb : A.xa;
    // input staging: thread tid fetches column tid of every channel's [l00 - 1, l00 + 256] window, threads 0 .. 2 Cx - 1 the two
    // remaining columns; the NEXT item's values travel in registers while the current item is computed
    float pf[CID_PF];
    const int xc = tid >> 1, xp = CID_ROWS + (tid & 1);
    auto fetch = [&](int item) {
        const int b = item / A.slices, l00 = (item - b * A.slices) * CID_ROWS;
        const float* xb = x + (size_t)b * Cx * L + l00 - 1;
        const int l = l00 + tid - 1;
        const bool ok = l >= 0 && l < L;
#pragma unroll
        for (int k = 0; k < CID_MAX_CX; ++k) pf[k] = (k < Cx && ok) ? xb[(size_t)k * L + tid] : 0.0f;
        pf[CID_MAX_CX] = (tid < 2 * Cx && l00 + xp - 1 < L) ? xb[(size_t)xc * L + xp] : 0.0f;
    };
    if ((int)blockIdx.x < A.n_items) fetch(blockIdx.x);
    const unsigned long long half2 = pkf2(0.5f, 0.5f);
    const float2* wsf = reinterpret_cast<const float2*>(ws);
    for (int item = blockIdx.x; item < A.n_items; item += gridDim.x) {
        const int b = item / A.slices, l00 = (item % A.slices) * CID_ROWS;
        __syncthreads();                 // the previous item's readers of xs / cfs are done
#pragma unroll
        for (int k = 0; k < CID_MAX_CX; ++k)
            if (k < Cx) xs[k * CID_XP + tid] = pf[k];
        if (tid < 2 * Cx) xs[xc * CID_XP + xp] = pf[CID_MAX_CX];
        {
            const float* cf = A.coef + (size_t)b * 32 * CF;
            for (int i = tid; i < 32 * CF; i += 256) cfs[i] = cf[i];
        }
        __syncthreads();
        // the next item's input travels while this one is computed
        if (item + (int)gridDim.x < A.n_items) fetch(item + gridDim.x);
#pragma unroll 1
        for (int rt = warp; rt < CID_ROWS / 16; rt += 8) {
            const int rbase = rt * 16;
            if (l00 + rbase >= L) break;
            float acc[8][4];                     // the conv WITHOUT its bias (folded into the B coefficient)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
            const float* x0 = xs + rbase + g;
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
                if (ks < KS) {
                    uint32_t af[4];
                    af[0] = cid_tf32(x0[aoff[ks][0]]);
                    af[1] = cid_tf32(x0[aoff[ks][0] + 8]);
                    af[2] = cid_tf32(x0[aoff[ks][1]]);
                    af[3] = cid_tf32(x0[aoff[ks][1] + 8]);
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        const float2 bf = wsf[(ks * 8 + nt) * 32 + lane];
                        cid_mma_tf32(acc[nt], af, __float_as_uint(bf.x), __float_as_uint(bf.y));
                    }
                }
            }
            // ---- epilogue on the fragments: rows rbase + g (acc[.][0..1]) and rbase + g + 8 (acc[.][2..3]), channels 8 nt + 2 t4 (+1)
            unsigned long long cv0[NCR], cv1[NCR];                      // cond = input channels 1 .. Cc at my two rows, as (v, v) pairs
#pragma unroll
            for (int jc = 0; jc < NCA; ++jc) {
                const int ro = jc < Cc ? (1 + jc) * CID_XP : Cx * CID_XP;   // weights beyond Cc are zero; read the zero row
                const float v0 = x0[ro + 1], v1 = x0[ro + 9];
                cv0[jc] = pkf2(v0, v0);
                cv1[jc] = pkf2(v1, v1);
            }
            const bool odd = (g & 1) != 0;
            uint32_t r0[8], r1[8], rp[8];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float* cf = cfs + (nt * 4 + t4) * CF;
                const ulonglong2 c_ab = *reinterpret_cast<const ulonglong2*>(cf);
                const ulonglong2 c_ge = *reinterpret_cast<const ulonglong2*>(cf + 4);
                const unsigned long long h0 = ffma2(pkf2(acc[nt][0], acc[nt][1]), c_ab.x, c_ab.y);
                const unsigned long long h1 = ffma2(pkf2(acc[nt][2], acc[nt][3]), c_ab.x, c_ab.y);
                float a0, a1, a2, a3, t0, t1, t2, t3;
                upk2(h0, a0, a1);
                upk2(h1, a2, a3);
                asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t2) : "f"(a2));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t3) : "f"(a3));
                unsigned long long o0 = ffma2(ffma2(h0, pkf2(t0, t1), h0), c_ge.x, c_ge.y);
                unsigned long long o1 = ffma2(ffma2(h1, pkf2(t2, t3), h1), c_ge.x, c_ge.y);
                if (CC != 0) {
#pragma unroll
                    for (int jc = 0; jc < NCA; ++jc) {
                        const unsigned long long wv = *reinterpret_cast<const unsigned long long*>(cf + 8 + 2 * jc);
                        o0 = ffma2(wv, cv0[jc], o0);
                        o1 = ffma2(wv, cv1[jc], o1);
                    }
                }
                float lo, hi;
                upk2(o0, lo, hi);
                r0[nt] = pack_bf16x2(lo, hi);
                upk2(o1, lo, hi);
                r1[nt] = pack_bf16x2(lo, hi);
                // 2:1 pooling: rows g and g ^ 1 sit in lanes t and t ^ 4.  The even lane averages the row pair of the upper half
                // (rows g, g+1), the odd lane the pair of the lower half (rows g+7, g+8): each sends the partner what it needs.
                const unsigned long long got = __shfl_xor_sync(0xffffffffu, odd ? o0 : o1, 4);
                upk2(fmul2(fadd2(odd ? o1 : o0, got), half2), lo, hi);
                rp[nt] = pack_bf16x2(lo, hi);
            }
            // ---- staging: out tile [16 rows][128 B], pooled tile [8 rows][128 B]; 16-byte chunk ch of row r at (ch ^ (r & 7))
            __syncwarp();                        // the previous tile's copy-out has finished reading the staging tiles
            {
                const int mrow = lane & 7, mi = lane >> 3;           // stmatrix: lane supplies the address of row mrow of matrix mi
#pragma unroll
                for (int c4 = 0; c4 < 2; ++c4) {
                    const int nt = c4 * 4 + mi;
                    cid_stmatrix_x4(stg + (uint32_t)mrow * 128u + (uint32_t)((nt ^ mrow) * 16), r0[c4 * 4], r0[c4 * 4 + 1], r0[c4 * 4 + 2],
                                    r0[c4 * 4 + 3]);
                    cid_stmatrix_x4(stg + (uint32_t)(8 + mrow) * 128u + (uint32_t)((nt ^ mrow) * 16), r1[c4 * 4], r1[c4 * 4 + 1],
                                    r1[c4 * 4 + 2], r1[c4 * 4 + 3]);
                }
                const int prow = odd ? 4 + (g >> 1) : (g >> 1);      // pooled row of this lane inside the tile
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(stg + 2048u + (uint32_t)prow * 128u + (uint32_t)((nt ^ prow) * 16) +
                                                                    (uint32_t)t4 * 4u),
                                 "r"(rp[nt])
                                 : "memory");
            }
            __syncwarp();
            // ---- coalesced copy-out: 128 chunks of 16 B (out), 64 chunks (pooled)
            {
                bf16* outb = A.out + ((size_t)b * L + l00 + rbase) * C;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int id = lane + 32 * i, r = id >> 3, ch = id & 7;
                    if (l00 + rbase + r < L) {
                        const uint4 v = *reinterpret_cast<const uint4*>(stg_g + r * 128 + ((ch ^ (r & 7)) * 16));
                        *reinterpret_cast<uint4*>(outb + (size_t)r * C + ch * 8) = v;
                    }
                }
                if (A.pooled != nullptr) {
                    bf16* poolb = A.pooled + ((size_t)b * (L / 2) + (l00 + rbase) / 2) * C;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int id = lane + 32 * i, r = id >> 3, ch = id & 7;
                        if (l00 + rbase + 2 * r + 1 < L) {
                            const uint4 v = *reinterpret_cast<const uint4*>(stg_g + 2048 + r * 128 + ((ch ^ (r & 7)) * 16));
                            *reinterpret_cast<uint4*>(poolb + (size_t)r * C + ch * 8) = v;
                        }
                    }
                }
            }
        }
    }
}

static int cid_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    }
    return n;
}

// floats of the coefficient workspace gw_conv_in_direct needs (0: the shape is not supported, use gw_conv_in_gn / gw_conv_in)
extern "C" long gw_conv_in_direct_ws_floats(int B, int Cx, int L, int C, int Cc) {
    if (C != CID_C || Cx < 1 || Cx > CID_MAX_CX || Cc < 0 || Cc > CID_MAX_CC || 1 + Cc > Cx || L < 4 || (L % 2) != 0 || B < 1) return 0;
    const int nca = (Cc == 0 || Cc == 1 || Cc == 5) ? Cc : CID_MAX_CC;
    return (long)B * 32 * CID_CF(nca);
}

extern "C" int gw_conv_in_direct(const float* x, const float* x_alt, const int* step_ptr, int B, int Cx, int L, const float* w,
                                 const float* bias, int C, const float* gn_w, const float* gn_b, int Cc, const float* wc,
                                 const float* bc, const float* film, int film_off, long film_b_stride, long film_step_stride,
                                 void* out, void* pooled, float* coef_ws, void* stream) {
    GW_REQUIRE(gw_conv_in_direct_ws_floats(B, Cx, L, C, Cc) > 0, "gw_conv_in_direct: unsupported shape (Cx=%d L=%d C=%d Cc=%d)", Cx, L, C, Cc);
    GW_REQUIRE(x != nullptr && w != nullptr && bias != nullptr && gn_w != nullptr && gn_b != nullptr && film != nullptr &&
                   out != nullptr && coef_ws != nullptr && (Cc == 0 || (wc != nullptr && bc != nullptr)),
               "gw_conv_in_direct: null pointer");
    CidArgs A;
    A.xa = x; A.xb = x_alt ? x_alt : x; A.step_ptr = step_ptr; A.w = w; A.bias = bias; A.gn_w = gn_w; A.gn_b = gn_b;
    A.wc = wc; A.bc = bc; A.film = film; A.coef = coef_ws; A.out = (bf16*)out; A.pooled = (bf16*)pooled;
    A.film_b_stride = film_b_stride; A.film_step_stride = film_step_stride; A.film_off = film_off;
    A.B = B; A.Cx = Cx; A.L = L; A.Cc = Cc;
    A.slices = (L + CID_ROWS - 1) / CID_ROWS;
    A.n_items = B * A.slices;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem_m = (size_t)Cx * (CID_CH + 8) * 4;
    const int nca = (Cc == 0 || Cc == 1 || Cc == 5) ? Cc : CID_MAX_CC;
    const size_t smem = (size_t)((Cx + 1) * CID_XP + 3 * 8 * 32 * 2 + 32 * CID_CF(nca)) * 4 + 8 * 3072;
    int grid = 2 * cid_sm_count();
    if (grid > A.n_items) grid = A.n_items;
#define CID_GO(CCV, NCAV)                                                                                                    \
    do {                                                                                                                     \
        GW_CUDA(cudaFuncSetAttribute(in_moments_kernel<NCAV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m));     \
        GW_CUDA(gw_launch_pdl(in_moments_kernel<NCAV>, dim3(B), dim3(256), smem_m, st, A));                                   \
        GW_CUDA(cudaFuncSetAttribute(conv_in_direct_kernel<CCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        GW_CUDA(gw_launch_pdl(conv_in_direct_kernel<CCV>, dim3(grid), dim3(256), smem, st, A));                               \
    } while (0)
    if (Cc == 0) CID_GO(0, 0);
    else if (Cc == 1) CID_GO(1, 1);
    else if (Cc == 5) CID_GO(5, 5);
    else CID_GO(-1, CID_MAX_CC);
#undef CID_GO
    GW_LAUNCH_CHECK();
    return GW_OK;
}
