// Fused FIRST block for inference (sm_100a): Conv1d(C_in -> 64, k=3) -> GroupNorm(8) -> SiLU -> + cond 1x1 conv -> FiLM ->
// out [B, L, 64] bf16 and avg_pool1d(out, 2), in one kernel, without the raw conv tensor (models.py:160-173, 188-193, 204-208).
//
// K = 3*C_in <= 21 is far too small for a tcgen05 tile pipeline, but the conv was the issue-bound part of this kernel on CUDA
// cores (576-1344 FMAs per row); for C_in <= 8 it now runs as warp-level tf32 MMAs (mma.sync.m16n8k8: a warp owns 16 rows x 64
// channels, K = 3*C_in padded to 16 / 24, A fragments straight from the staged fp32 input, B fragments from shared memory),
// fp32 accumulation.  What the kernel removes is the 134 MB raw write + read between gw_conv_in and gw_gn_apply.  Like conv_gn.cuh, a GROUP of G = L/256
// (L/512 beyond L = 8192) persistent CTAs owns a sample: each CTA computes its rows ONCE into a bf16 slice in shared memory,
// sums the GroupNorm statistics, and exchanges them with the group through {value, epoch} packets (xchg.cuh).  The exchange
// latency is hidden by software pipelining: a CTA convolves sample s+1 into the second slice (and fetches the input of s+2
// into registers) before it waits for the statistics of sample s and normalises / activates / modulates slice s straight
// out of shared memory; two CTAs share an SM, so one's conv overlaps the other's apply phase.
// The numerics are those of the unfused pair: statistics of the bf16-rounded conv output, apply from the bf16 values.
#include "common.cuh"
#include "../../include/gwb200.h"
#include "xchg.cuh"

// rows of a sample per CTA: 256 (two CTAs per SM: one convolves while the other normalises) up to L = 8192, else 512
#define CIG_C 64
#define CIG_MAX_CC 8
#define CIG_PF 8                       // prefetch registers per thread (covers C_in <= 7 at 256 rows)

struct CigArgs {
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
    bf16* out;
    bf16* pooled;
    void* sync;
    long film_b_stride, film_step_stride;
    int film_off, B, Cx, L, Cc, G, n_groups;
};

// weight staging: [Cx*3][C] for the CUDA-core conv, or tf32 B fragments [KS][8 n-tiles][32 lanes][2] for the MMA conv
__host__ __device__ inline int cig_ksteps(int Cx) { return Cx <= 8 ? (3 * Cx + 7) / 8 : 0; }     // 0: CUDA-core conv
__host__ __device__ inline int cig_ws_floats(int Cx) {
    const int a = Cx * 3 * CIG_C, b = cig_ksteps(Cx) * 8 * 32 * 2;
    return a > b ? a : b;
}
__device__ __forceinline__ uint32_t cig_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void cig_mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

static __device__ __noinline__ void cig_timeout(int b, int src) {
    printf("gwb200 conv_in_gn kernel: statistics exchange timed out (block %d sample %d source %d)\n", blockIdx.x, b, src);
    __trap();
}

template <int CC, int CIG_ROWS>      // cond channels: 1, 5, or -1 (any 0..8, zero-padded); rows per CTA
__global__ void __launch_bounds__(256, CIG_ROWS == 256 ? 2 : 1) conv_in_gn_kernel(const CigArgs A) {
    constexpr int NCA = CC > 0 ? CC : CIG_MAX_CC;
    constexpr int CIG_XP = CIG_ROWS + 8;          // row pitch of the staged input (floats)
    constexpr int C = CIG_C;
    extern __shared__ __align__(16) unsigned char smem[];
    bf16* slice = reinterpret_cast<bf16*>(smem);                            // [2][ROWS][64 ch]
    float* xs = reinterpret_cast<float*>(smem + 2 * CIG_ROWS * C * 2);      // [2][Cx][XP]: xs[c][j] = x[c][l00 - 1 + j]
    float* ws = xs + 2 * A.Cx * CIG_XP;                                     // [Cx*3][C], split-octet layout
    float* bs = ws + cig_ws_floats(A.Cx);                                   // [C]
    float* wst = bs + C;                                                    // [8 warps][8 octets][2]
    float* s_x = wst + 128;                                                 // [G][16]
    float* s_mr = s_x + XCHG_MAX_G * 16;                                    // [8 groups][mean, rstd]
    const int Cx = A.Cx, L = A.L, Cc = A.Cc, G = A.G;
    const int grp = blockIdx.x / G, j_cta = blockIdx.x % G;
    const int l00 = j_cta * CIG_ROWS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int oct = lane & 7;                            // conv: my 8 output channels = GroupNorm group `oct`
    const int pg = warp * 4 + (lane >> 3);               // conv: position group (4 positions) inside a 128-position block
    unsigned int* ctrl = reinterpret_cast<unsigned int*>(A.sync);

    // weights as [ck][half][octet][4]: the 8 octet lanes of a quarter-warp read 8 consecutive float4 (no bank conflicts)
    auto split = [](int co) { return ((co >> 2) & 1) * (C / 2) + (co >> 3) * 4 + (co & 3); };
    const int KS = cig_ksteps(Cx);                       // > 0: tf32 MMA conv
    if (KS > 0) {
        // B fragment of (k-step ks, n-tile nt) for lane (g = lane>>2, t = lane&3): W[kk = 8 ks + t (+4)][co = 8 nt + g]
        for (int i = tid; i < KS * 8 * 32 * 2; i += 256) {
            const int j = i & 1, ln = (i >> 1) & 31, nt = (i >> 6) & 7, ks = i >> 9;
            const int kk = ks * 8 + (ln & 3) + 4 * j, co = nt * 8 + (ln >> 2);
            const float v = kk < 3 * Cx ? A.w[(size_t)co * Cx * 3 + kk] : 0.0f;
            ws[i] = __uint_as_float(cig_tf32(v));
        }
    } else {
        for (int i = tid; i < Cx * 3 * C; i += 256) {
            const int co = i % C, ck = i / C;
            ws[ck * C + split(co)] = A.w[(size_t)co * Cx * 3 + ck];
        }
    }
    // A fragment columns of this lane: kk = 8 ks + t (+4) -> (input channel, tap) -> offset into the staged input
    int aoff[3][2];
#pragma unroll
    for (int ks = 0; ks < 3; ++ks)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int kk = ks * 8 + (lane & 3) + 4 * j;
            aoff[ks][j] = kk < 3 * Cx ? (kk / 3) * CIG_XP + kk % 3 : -1;
        }
    for (int i = tid; i < C; i += 256) bs[i] = A.bias[i];
    // apply phase: this thread always owns channels 8*ao .. 8*ao+7 (one GroupNorm group) of row pairs (tid>>3) + 32*jj
    const int ao = tid & 7;
    const double inv_n = 1.0 / (8.0 * (double)L);
    __syncthreads();
    // parameters only so far; the input, the exchange buffer and the step counter belong to the previous kernel (PDL, common.cuh)
    pdl_wait();
    pdl_launch_dependents();
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(ctrl) + 1u;
    const int step = A.step_ptr != nullptr ? *A.step_ptr : 0;
    const float* x = (step & 1) ? A.xb : A.xa;

    // input staging: x[b, c, l00-1 .. l00+ROWS] -> xs; the NEXT sample's values are fetched into registers before phase 2 so
    // that their HBM latency is hidden behind it
    auto load_x = [&](int b, int i) {
        const int c = i / CIG_XP, p = i % CIG_XP;
        const int l = l00 + p - 1;
        return (p < CIG_ROWS + 2 && l >= 0 && l < L) ? x[((size_t)b * Cx + c) * L + l] : 0.0f;
    };
    float pf[CIG_PF];
    bool pf_valid = false;
    const bool can_pf = Cx * CIG_XP <= CIG_PF * 256;

    for (int it = 0;; ++it) {
        const int b = grp + it * A.n_groups;
        const bool has = b < A.B;
        if (has) {
            // ================= phase 1: conv of sample b into slice[it & 1], statistics, publish =================
            float* xsb = xs + (it & 1) * Cx * CIG_XP;
            bf16* sl = slice + (size_t)(it & 1) * CIG_ROWS * C;
            if (!pf_valid) {
                for (int i = tid; i < Cx * CIG_XP; i += 256) xsb[i] = load_x(b, i);
            } else {
#pragma unroll
                for (int k = 0; k < CIG_PF; ++k)
                    if (tid + k * 256 < Cx * CIG_XP) xsb[tid + k * 256] = pf[k];
            }
            __syncthreads();
            if (KS > 0) {
                // ---- tf32 MMA conv: warp = 16 rows x 64 channels per tile; n-tile nt = GroupNorm group nt ----
                const int g = lane >> 2, t4 = lane & 3;
                float s1v[8], s2v[8];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) s1v[nt] = s2v[nt] = 0.0f;
                const float2* wsf = reinterpret_cast<const float2*>(ws);
                uint32_t* sl32 = reinterpret_cast<uint32_t*>(sl);
#pragma unroll 1
                for (int rt = warp; rt < CIG_ROWS / 16; rt += 8) {
                    const int rbase = rt * 16;
                    if (l00 + rbase >= L) break;
                    float acc[8][4];
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        const float2 b2 = *reinterpret_cast<const float2*>(bs + nt * 8 + 2 * t4);
                        acc[nt][0] = b2.x; acc[nt][1] = b2.y; acc[nt][2] = b2.x; acc[nt][3] = b2.y;
                    }
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks) {
                        if (ks < KS) {
                            uint32_t af[4];
                            const float* x0 = xsb + rbase + g;
                            af[0] = aoff[ks][0] >= 0 ? cig_tf32(x0[aoff[ks][0]]) : 0u;
                            af[1] = aoff[ks][0] >= 0 ? cig_tf32(x0[aoff[ks][0] + 8]) : 0u;
                            af[2] = aoff[ks][1] >= 0 ? cig_tf32(x0[aoff[ks][1]]) : 0u;
                            af[3] = aoff[ks][1] >= 0 ? cig_tf32(x0[aoff[ks][1] + 8]) : 0u;
#pragma unroll
                            for (int nt = 0; nt < 8; ++nt) {
                                const float2 bf = wsf[(ks * 8 + nt) * 32 + lane];
                                cig_mma_tf32(acc[nt], af, __float_as_uint(bf.x), __float_as_uint(bf.y));
                            }
                        }
                    }
                    // rows rbase + g and rbase + g + 8, channels 8 nt + 2 t4 (+1): bf16 into the swizzled slice, statistics
                    const bool v0 = l00 + rbase + g < L, v1 = l00 + rbase + g + 8 < L;
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        const uint32_t w0 = pack_bf16x2(acc[nt][0], acc[nt][1]), w1 = pack_bf16x2(acc[nt][2], acc[nt][3]);
                        const int col = ((nt ^ g) * 8 + 2 * t4) >> 1;
                        sl32[(size_t)(rbase + g) * (C / 2) + col] = w0;
                        sl32[(size_t)(rbase + g + 8) * (C / 2) + col] = w1;
                        if (v0) {
                            const float lo = __uint_as_float(w0 << 16), hi = __uint_as_float(w0 & 0xffff0000u);
                            s1v[nt] += lo + hi;
                            s2v[nt] = fmaf(lo, lo, fmaf(hi, hi, s2v[nt]));
                        }
                        if (v1) {
                            const float lo = __uint_as_float(w1 << 16), hi = __uint_as_float(w1 & 0xffff0000u);
                            s1v[nt] += lo + hi;
                            s2v[nt] = fmaf(lo, lo, fmaf(hi, hi, s2v[nt]));
                        }
                    }
                }
                float sv16[16];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) { sv16[nt] = s1v[nt]; sv16[8 + nt] = s2v[nt]; }
                const float tot = warp_reduce_multi<16>(sv16, lane);      // value j lands in lanes 2j, 2j+1
                if ((lane & 1) == 0) wst[(warp * 8 + ((lane >> 1) & 7)) * 2 + (lane >> 4)] = tot;
            } else {
                float s1 = 0.0f, s2 = 0.0f;
    #pragma unroll 1
                for (int blk = 0; blk < CIG_ROWS / 128; ++blk) {
                    const int r0 = blk * 128 + pg * 4;               // first of my 4 rows inside the CTA's slice
                    if (l00 + blk * 128 >= L) break;
                    float acc[4][8];
    #pragma unroll
                    for (int u = 0; u < 4; ++u)
    #pragma unroll
                        for (int j = 0; j < 8; ++j) acc[u][j] = bs[oct * 8 + j];
                    for (int ci = 0; ci < Cx; ++ci) {
                        const float* xr = xsb + ci * CIG_XP + r0;
                        const float4 xa4 = *reinterpret_cast<const float4*>(xr);
                        const float2 xb2 = *reinterpret_cast<const float2*>(xr + 4);
                        const float xv[6] = {xa4.x, xa4.y, xa4.z, xa4.w, xb2.x, xb2.y};
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
    #pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint4 pk;
                        pk.x = pack_bf16x2(acc[u][0], acc[u][1]);
                        pk.y = pack_bf16x2(acc[u][2], acc[u][3]);
                        pk.z = pack_bf16x2(acc[u][4], acc[u][5]);
                        pk.w = pack_bf16x2(acc[u][6], acc[u][7]);
                        *reinterpret_cast<uint4*>(sl + (size_t)(r0 + u) * C + ((oct ^ ((r0 + u) & 7)) * 8)) = pk;
                        if (l00 + r0 + u < L) {
                            const uint32_t wds[4] = {pk.x, pk.y, pk.z, pk.w};
    #pragma unroll
                            for (int q = 0; q < 4; ++q) {           // statistics of the values as stored (bf16)
                                const float lo = __uint_as_float(wds[q] << 16), hi = __uint_as_float(wds[q] & 0xffff0000u);
                                s1 += lo + hi;
                                s2 = fmaf(lo, lo, fmaf(hi, hi, s2));
                            }
                        }
                    }
                }
                // fold the 4 position groups of the warp that share an octet (lanes differing in bits 3, 4), then the 8 warps
                s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
                s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
                s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
                if (lane < 8) {
                    wst[(warp * 8 + oct) * 2 + 0] = s1;
                    wst[(warp * 8 + oct) * 2 + 1] = s2;
                }
        }
            __syncthreads();
            if (tid < 16) {
                float v = 0.0f;
#pragma unroll
                for (int wi = 0; wi < 8; ++wi) v += wst[(wi * 8 + (tid >> 1)) * 2 + (tid & 1)];
                st_relaxed_u64(xchg_slot(A.sync, b, j_cta) + tid,
                               ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v));
            }
        }
        pf_valid = false;
        if (can_pf && b + A.n_groups < A.B) {
#pragma unroll
            for (int k = 0; k < CIG_PF; ++k) pf[k] = tid + k * 256 < Cx * CIG_XP ? load_x(b + A.n_groups, tid + k * 256) : 0.0f;
            pf_valid = true;
        }
        if (it > 0) {
            // ================= phase 2: statistics of the previous sample -> apply its slice -> out, pooled =================
            const int bp = b - A.n_groups;
            const float* xsb = xs + ((it - 1) & 1) * Cx * CIG_XP;
            const bf16* sl = slice + (size_t)((it - 1) & 1) * CIG_ROWS * C;
            // FiLM row of my channels (issued before the poll so that the loads travel meanwhile)
            const float* fr = A.film + (size_t)step * A.film_step_stride + (size_t)bp * A.film_b_stride + A.film_off;
            float fg[8], fb[8], gw[8], gb[8], bcv[8], wcv[NCA][8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = ao * 8 + j;
                fg[j] = fr[c];
                fb[j] = fr[C + c];
                gw[j] = A.gn_w[c];                       // re-read per sample (L1 hits): not live during the conv phase
                gb[j] = A.gn_b[c];
                bcv[j] = Cc > 0 ? A.bc[c] : 0.0f;
#pragma unroll
                for (int jc = 0; jc < NCA; ++jc) wcv[jc][j] = jc < Cc ? A.wc[c * Cc + jc] : 0.0f;
            }
            for (int i = tid; i < G * 16; i += 256) {
                const unsigned long long* src = xchg_slot(A.sync, bp, i >> 4) + (i & 15);
                unsigned long long pk = ld_relaxed_u64(src);
                if ((unsigned int)(pk >> 32) != epoch) {
                    const long long t0 = clock64();
                    do {
                        pk = ld_relaxed_u64(src);
                        if (clock64() - t0 > 4000000000LL) cig_timeout(bp, i >> 4);
                    } while ((unsigned int)(pk >> 32) != epoch);
                }
                s_x[i] = __uint_as_float((unsigned int)pk);
            }
            __syncthreads();
            // one thread per GroupNorm group folds the G packets in fp64 (every thread doing it was ~5 % of the kernel's instructions)
            if (tid < 8) {
                double a1 = 0.0, a2 = 0.0;
                for (int s = 0; s < G; ++s) {
                    a1 += (double)s_x[s * 16 + tid * 2];
                    a2 += (double)s_x[s * 16 + tid * 2 + 1];
                }
                const double mean = a1 * inv_n;
                double var = a2 * inv_n - mean * mean;
                if (var < 0.0) var = 0.0;
                s_mr[tid * 2 + 0] = (float)mean;
                s_mr[tid * 2 + 1] = (float)(1.0 / sqrt(var + 1e-5));
            }
            __syncthreads();
            const float meanf = s_mr[ao * 2 + 0], rstd = s_mr[ao * 2 + 1];
            // coefficients of my 8 channels, packed in pairs: h = A x + B (half the GroupNorm output), out = silu G + E + sum W c
            f32x2 hA[4], hB[4], Gp[4], Ep[4], Wp[NCA][4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float a_[2], b_[2], g_[2], e_[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = 2 * q + h;
                    const float a = rstd * gw[j];
                    a_[h] = 0.5f * a;
                    b_[h] = 0.5f * (gb[j] - meanf * a);
                    g_[h] = 1.0f + fg[j];
                    e_[h] = fmaf(bcv[j], g_[h], fb[j]);
                }
                hA[q] = pkf2(a_[0], a_[1]);
                hB[q] = pkf2(b_[0], b_[1]);
                Gp[q] = pkf2(g_[0], g_[1]);
                Ep[q] = pkf2(e_[0], e_[1]);
#pragma unroll
                for (int jc = 0; jc < NCA; ++jc) Wp[jc][q] = pkf2(wcv[jc][2 * q] * g_[0], wcv[jc][2 * q + 1] * g_[1]);
            }
            const f32x2 half2 = pkf2(0.5f, 0.5f);
            bf16* outb = A.out + ((size_t)bp * L + l00) * C + ao * 8;
            bf16* poolb = A.pooled != nullptr ? A.pooled + ((size_t)bp * (L / 2) + l00 / 2) * C + ao * 8 : nullptr;
#pragma unroll 2
            for (int jj = 0; jj < CIG_ROWS / 64; ++jj) {
                const int p = (tid >> 3) + 32 * jj;              // row pair: rows 2p, 2p+1 of this CTA
                if (l00 + 2 * p >= L) break;
                f32x2 o[2][4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = 2 * p + h;
                    const uint4 xr = *reinterpret_cast<const uint4*>(sl + (size_t)r * C + ((ao ^ (r & 7)) * 8));
                    const uint32_t wds[4] = {xr.x, xr.y, xr.z, xr.w};
                    float cv[NCA];
#pragma unroll
                    for (int jc = 0; jc < NCA; ++jc) cv[jc] = jc < Cc ? xsb[(1 + jc) * CIG_XP + r + 1] : 0.0f;   // cond = input channels
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const f32x2 hh = ffma2(pk2(wds[q] << 16, wds[q] & 0xffff0000u), hA[q], hB[q]);
                        float h0, h1, t0, t1;
                        upk2(hh, h0, h1);
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
                        f32x2 v = ffma2(ffma2(hh, pkf2(t0, t1), hh), Gp[q], Ep[q]);
#pragma unroll
                        for (int jc = 0; jc < NCA; ++jc) v = ffma2(Wp[jc][q], pkf2(cv[jc], cv[jc]), v);
                        o[h][q] = v;
                    }
                    if (l00 + r < L) {
                        uint4 pk;
                        float a0, a1f;
                        upk2(o[h][0], a0, a1f); pk.x = pack_bf16x2(a0, a1f);
                        upk2(o[h][1], a0, a1f); pk.y = pack_bf16x2(a0, a1f);
                        upk2(o[h][2], a0, a1f); pk.z = pack_bf16x2(a0, a1f);
                        upk2(o[h][3], a0, a1f); pk.w = pack_bf16x2(a0, a1f);
                        *reinterpret_cast<uint4*>(outb + (size_t)r * C) = pk;
                    }
                }
                if (poolb != nullptr && l00 + 2 * p + 1 < L) {
                    uint4 pk;
                    float a0, a1f;
                    upk2(fmul2(fadd2(o[0][0], o[1][0]), half2), a0, a1f); pk.x = pack_bf16x2(a0, a1f);
                    upk2(fmul2(fadd2(o[0][1], o[1][1]), half2), a0, a1f); pk.y = pack_bf16x2(a0, a1f);
                    upk2(fmul2(fadd2(o[0][2], o[1][2]), half2), a0, a1f); pk.z = pack_bf16x2(a0, a1f);
                    upk2(fmul2(fadd2(o[0][3], o[1][3]), half2), a0, a1f); pk.w = pack_bf16x2(a0, a1f);
                    *reinterpret_cast<uint4*>(poolb + (size_t)p * C) = pk;
                }
            }
        }
        if (!has) break;
        __syncthreads();          // wst / s_x / the slice two samples back are reused by the next iteration
    }
    __syncthreads();
    if (tid == 0) xchg_finish(ctrl);
}

static int cig_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    }
    return n;
}

// 0 if the fused first block cannot run this shape (then use gw_conv_in + gw_gn_apply), else the CTAs per sample
extern "C" int gw_conv_in_gn_group(int Cx, int L, int C, int Cc) {
    if (C != CIG_C || Cx < 1 || Cx > 16 || Cc < 0 || Cc > CIG_MAX_CC || 1 + Cc > Cx || L < 2 || (L % 2) != 0) return 0;
    const int rows = L <= 8192 ? 256 : 512;
    const int G = (L + rows - 1) / rows;
    if (G > XCHG_MAX_G || G > cig_sm_count()) return 0;
    return G;
}

extern "C" int gw_conv_in_gn(const float* x, const float* x_alt, const int* step_ptr, int B, int Cx, int L, const float* w,
                             const float* bias, int C, const float* gn_w, const float* gn_b, int Cc, const float* wc,
                             const float* bc, const float* film, int film_off, long film_b_stride, long film_step_stride,
                             void* out, void* pooled, void* sync, void* stream) {
    const int G = gw_conv_in_gn_group(Cx, L, C, Cc);
    GW_REQUIRE(G > 0, "gw_conv_in_gn: unsupported shape (Cx=%d L=%d C=%d Cc=%d)", Cx, L, C, Cc);
    GW_REQUIRE(x != nullptr && w != nullptr && bias != nullptr && gn_w != nullptr && gn_b != nullptr && film != nullptr &&
                   out != nullptr && sync != nullptr && (Cc == 0 || (wc != nullptr && bc != nullptr)),
               "gw_conv_in_gn: null pointer");
    CigArgs A;
    A.xa = x; A.xb = x_alt ? x_alt : x; A.step_ptr = step_ptr; A.w = w; A.bias = bias; A.gn_w = gn_w; A.gn_b = gn_b;
    A.wc = wc; A.bc = bc; A.film = film; A.out = (bf16*)out; A.pooled = (bf16*)pooled; A.sync = sync;
    A.film_b_stride = film_b_stride; A.film_step_stride = film_step_stride; A.film_off = film_off;
    A.B = B; A.Cx = Cx; A.L = L; A.Cc = Cc; A.G = G;
    const int rows = L <= 8192 ? 256 : 512;
    const size_t smem = (size_t)2 * rows * CIG_C * 2 + (size_t)(2 * Cx * (rows + 8) + cig_ws_floats(Cx) + CIG_C + 128 + XCHG_MAX_G * 16 + 16) * 4;
    GW_REQUIRE(smem <= 232448, "gw_conv_in_gn: shared memory %zu too large", smem);
    cudaStream_t st = (cudaStream_t)stream;
    // every CTA of a group spins on its peers: the grid must be co-resident -> size it from the occupancy of this variant
#define CIG_GO(CCV, ROWSV)                                                                                               \
    do {                                                                                                                 \
        GW_CUDA(cudaFuncSetAttribute(conv_in_gn_kernel<CCV, ROWSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        int occ = 0;                                                                                                     \
        GW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_in_gn_kernel<CCV, ROWSV>, 256, smem));           \
        GW_REQUIRE(occ >= 1, "gw_conv_in_gn: kernel does not fit on an SM");                                              \
        if (occ > 2) occ = 2;                                                                                            \
        A.n_groups = occ * cig_sm_count() / G;                                                                           \
        if (A.n_groups > B) A.n_groups = B;                                                                              \
        GW_REQUIRE(A.n_groups >= 1, "gw_conv_in_gn: a sample needs more CTAs than fit on the GPU");                      \
        GW_CUDA(gw_launch_pdl(conv_in_gn_kernel<CCV, ROWSV>, dim3(G * A.n_groups), dim3(256), smem, st, A));             \
    } while (0)
#define CIG_ROWS_GO(CCV)                      \
    do {                                      \
        if (rows == 256) CIG_GO(CCV, 256);    \
        else CIG_GO(CCV, 512);                \
    } while (0)
    if (Cc == 1) CIG_ROWS_GO(1);
    else if (Cc == 5) CIG_ROWS_GO(5);
    else CIG_ROWS_GO(-1);
#undef CIG_ROWS_GO
#undef CIG_GO
    GW_LAUNCH_CHECK();
    return GW_OK;
}
