// One-pass GroupNorm/SiLU/cond/FiLM backward for bf16 channels-last activations (sm_100a).
//
// The two-pass kernels (stream_gn.cu) read raw / d_out once for the per-sample sums and a second time to form d_raw:
// 5-6 tensor passes over HBM per layer.  Here a sample's operands are read ONCE: a group of G persistent CTAs owns a sample,
// every CTA pulls its row slice (raw, d_out, d_pooled, cond: <= ~44 KB) into shared memory with 1-D bulk copies, sums its part
// of the per-channel statistics from shared memory, publishes the two GroupNorm group sums through {value, epoch} packets
// (xchg.cuh, the protocol of conv_gn.cuh), and once all G packets of the sample are in, forms d_raw out of the same
// shared-memory slice (in place over raw) and bulk-stores it.  HBM traffic per layer: read raw + d_out (+ d_pooled/2 + cond),
// write d_raw = 3-3.5 tensor passes.  Four CTAs per SM are at different phases (load / sums / exchange / apply / store), which
// keeps loads in flight while others wait for their peers.
//
// Math: gn_bwd_stats_stream_kernel / gn_bwd_apply_stream_kernel (stream_gn.cu), i.e. the backward of models.py:160-173.
// The per-channel partial sums go to `partial` in the layout gn_bwd_finalize_kernel (backward.cu) expects, with n_rc = G.
#include "common.cuh"
#include "../../include/gwb200.h"
#include "tc_common.cuh"
#include "xchg.cuh"
#include "gn_bwd.cuh"

#define GBF_NCH 4            // chunks (mbarriers) per slice: the sums start while the tail of the slice is still in flight
#define GBF_MAX_CC 8
#define GBF_VB 3             // statistics folded per pass through the shared-memory reduction buffer

int g_gn_bwd_fused = 1;
int g_gn_bwd_fused_slice = 45056;      // largest slice (bytes of shared memory operands per CTA)

struct GbfPlan {
    int G, R;                                  // CTAs per sample, rows per CTA
    uint32_t off_do, off_pool, off_cond, off_red, off_x, off_pub, off_bar, smem;
};

static bool gbf_plan(int L, int C, int Cc, bool has_do, bool has_pool, GbfPlan* pl) {
    if (!(C == 64 || C == 128 || C == 256) || Cc < 0 || Cc > GBF_MAX_CC || L < 16) return false;
    const long row_bytes = (long)C * 2 * (1 + (has_do ? 1 : 0)) + (has_pool ? C : 0) + (long)Cc * 4;
    for (int G = 1; G <= XCHG_MAX_G; G *= 2) {
        if (L % (G * 4 * GBF_NCH) != 0) return false;                   // chunks of a multiple of 4 rows (16-byte bulk copies)
        const int R = L / G;
        if ((long)R * row_bytes > g_gn_bwd_fused_slice) continue;
        uint32_t o = (uint32_t)R * C * 2;
        pl->G = G; pl->R = R;
        pl->off_do = o;   if (has_do) o += (uint32_t)R * C * 2;
        pl->off_pool = o; if (has_pool) o += (uint32_t)R * C;
        pl->off_cond = o; o += (uint32_t)((R * Cc * 4 + 127) & ~127);
        pl->off_red = o;  o += (uint32_t)(256 / (C / 2)) * C * GBF_VB * 4;
        pl->off_x = o;    o += XCHG_MAX_G * 16 * 4;
        pl->off_pub = o;  o += 64;
        pl->off_bar = o;  o += GBF_NCH * 8;
        pl->smem = o;
        return true;
    }
    return false;
}

static __device__ __noinline__ void gbf_timeout(int b, int src) {
    printf("gwb200 gn_bwd_fused kernel: statistics exchange timed out (block %d sample %d source %d)\n", blockIdx.x, b, src);
    __trap();
}

template <int CC>      // cond channels: 0, 1, 5, or -1 (any <= 8)
__global__ void __launch_bounds__(256, 4) gn_bwd_fused_kernel(const GnBwdArgs a, const GbfPlan pl, int B, int n_groups,
                                                              float* __restrict__ partial, float* __restrict__ partial_bias,
                                                              bf16* __restrict__ d_raw, void* sync) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int NC = CC >= 0 ? CC : GBF_MAX_CC;
    constexpr int NV = 4 + NC;
    extern __shared__ __align__(128) uint8_t smem[];
    const int Cc = CC >= 0 ? CC : a.Cc, nvr = 4 + Cc;
    const int C = a.C, L = a.L, G = pl.G, R = pl.R, RC = R / GBF_NCH;
    const int grp = blockIdx.x / G, j_cta = blockIdx.x % G;
    const int tid = threadIdx.x;
    const int n_pair = C / 2, n_tr = 256 / n_pair;
    const int pr = tid % n_pair, tr = tid / n_pair;
    const int r0 = j_cta * R;
    const bool has_do = a.do_a != nullptr, has_pool = a.do_pool != nullptr;
    uint32_t* s_raw = reinterpret_cast<uint32_t*>(smem);                       // [R][C/2] channel pairs (d_raw is formed in place)
    const uint32_t* s_do = reinterpret_cast<const uint32_t*>(smem + pl.off_do);
    const uint32_t* s_pool = reinterpret_cast<const uint32_t*>(smem + pl.off_pool);
    const float* s_cond = reinterpret_cast<const float*>(smem + pl.off_cond);  // [R][Cc]
    float* red = reinterpret_cast<float*>(smem + pl.off_red);                  // [n_tr][C][GBF_VB]
    float* s_x = reinterpret_cast<float*>(smem + pl.off_x);                    // [G][16]
    float* s_pub = reinterpret_cast<float*>(smem + pl.off_pub);                // [16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.off_bar);
    unsigned int* ctrl = reinterpret_cast<unsigned int*>(sync);
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(ctrl) + 1u;
    if (tid == 0) {
        for (int k = 0; k < GBF_NCH; ++k) mbar_init(smem_u32(bars + k), 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    const int cg = C / 8, g = (2 * pr) / cg;
    const float gw0 = a.gn_w[2 * pr], gw1 = a.gn_w[2 * pr + 1], gb0 = a.gn_b[2 * pr], gb1 = a.gn_b[2 * pr + 1];
    const float gw_c = tid < C ? a.gn_w[tid] : 0.0f;
    const f32x2 GW2 = pkf2(gw0, gw1);
    const f32x2 half2 = pkf2(0.5f, 0.5f);
    const float inv_n = 1.0f / ((float)cg * (float)L);
    __syncthreads();

    for (int it = 0;; ++it) {
        const int b = grp + it * n_groups;
        if (b >= B) break;
        const uint32_t par = (uint32_t)(it & 1);
        if (tid == 0) {
            tma_wait_read<0>();                       // the previous sample's d_raw store has left the raw region
            const size_t row = (size_t)b * L + r0;
            for (int k = 0; k < GBF_NCH; ++k) {
                const uint32_t bar = smem_u32(bars + k);
                const uint32_t nb = (uint32_t)RC * C * 2;
                uint32_t total = nb;
                if (has_do) total += nb;
                if (has_pool) total += nb / 2;
                if (Cc > 0) total += (uint32_t)RC * Cc * 4;
                mbar_expect_tx(bar, total);
                const size_t rk = row + (size_t)k * RC;
                bulk_load(smem_u32(smem) + k * nb, (const bf16*)a.raw + rk * C, nb, bar);
                if (has_do) bulk_load(smem_u32(smem) + pl.off_do + k * nb, (const bf16*)a.do_a + rk * C, nb, bar);
                if (has_pool)
                    bulk_load(smem_u32(smem) + pl.off_pool + k * (nb / 2),
                              (const bf16*)a.do_pool + ((size_t)b * (L / 2) + ((r0 + k * RC) >> 1)) * C, nb / 2, bar);
                if (Cc > 0) bulk_load(smem_u32(smem) + pl.off_cond + k * RC * Cc * 4, a.cond + rk * Cc, (uint32_t)RC * Cc * 4, bar);
            }
        }
        // per-sample coefficients of my channel pair (the loads travel while the slice does)
        const float mean = a.stats[((size_t)b * 8 + g) * 2 + 0];
        const float rstd = a.stats[((size_t)b * 8 + g) * 2 + 1];
        const float* fr = a.film + (size_t)b * a.film_b_stride + a.film_off;
        const float fg0 = fr[2 * pr], fg1 = fr[2 * pr + 1];
        const f32x2 rs2 = pkf2(rstd, rstd), xo2 = pkf2(-mean * rstd, -mean * rstd);
        const float aa0 = rstd * gw0, aa1 = rstd * gw1;
        const f32x2 hA = pkf2(0.5f * aa0, 0.5f * aa1);
        const f32x2 hB = pkf2(0.5f * (gb0 - mean * aa0), 0.5f * (gb1 - mean * aa1));
        const f32x2 Gm = pkf2(1.0f + fg0, 1.0f + fg1);

        // ---------------- per-channel sums over my rows ----------------
        f32x2 acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = 0ull;
        for (int k = 0; k < GBF_NCH; ++k) {
            mbar_wait(smem_u32(bars + k), par);
            const int rend = (k + 1) * RC;
            for (int rb = k * RC + tr; rb < rend; rb += 2 * n_tr) {
#pragma unroll
                for (int uu = 0; uu < 2; ++uu) {
                    const int r = rb + uu * n_tr;
                    if (r >= rend) break;
                    const uint32_t xw = s_raw[r * n_pair + pr];
                    f32x2 dv = has_do ? bf2_lo(s_do[r * n_pair + pr]) : 0ull;
                    if (has_pool) dv = ffma2(bf2_lo(s_pool[(r >> 1) * n_pair + pr]), half2, dv);
                    const f32x2 x = bf2_lo(xw);
                    f32x2 z, act, dact;
                    sg_silu_pair(x, hA, hB, z, act, dact);
                    const f32x2 dn = fmul2(fmul2(dv, Gm), dact);
                    const f32x2 xh = ffma2(x, rs2, xo2);
                    acc[0] = fadd2(acc[0], dv);
                    acc[1] = ffma2(dv, act, acc[1]);
                    acc[2] = fadd2(acc[2], dn);
                    acc[3] = ffma2(dn, xh, acc[3]);
#pragma unroll
                    for (int jc = 0; jc < NC; ++jc) {
                        const float cvj = jc < Cc ? s_cond[r * Cc + jc] : 0.0f;
                        acc[4 + jc] = ffma2(dv, pkf2(cvj, cvj), acc[4 + jc]);
                    }
                }
            }
        }
        // fold the n_tr row lanes of every channel in fixed order, GBF_VB statistics per pass
        float sv[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) sv[v] = 0.0f;
#pragma unroll
        for (int vb = 0; vb < NV; vb += GBF_VB) {
            if (vb < nvr) {
#pragma unroll
                for (int q = 0; q < GBF_VB; ++q)
                    if (vb + q < NV) {
                        float lo, hi;
                        upk2(acc[vb + q], lo, hi);
                        red[((size_t)tr * C + 2 * pr) * GBF_VB + q] = lo;
                        red[((size_t)tr * C + 2 * pr + 1) * GBF_VB + q] = hi;
                    }
            }
            __syncthreads();
            if (tid < C && vb < nvr) {
#pragma unroll
                for (int q = 0; q < GBF_VB; ++q)
                    if (vb + q < NV) {
                        float s = 0.0f;
                        for (int t = 0; t < n_tr; ++t) s += red[((size_t)t * C + tid) * GBF_VB + q];
                        sv[vb + q] = s;
                    }
            }
            __syncthreads();
        }
        if (tid < C) {
            float* pt = partial + (((size_t)b * G + j_cta) * C + tid) * nvr;
#pragma unroll
            for (int v = 0; v < NV; ++v)
                if (v < nvr) pt[v] = sv[v];
            // GroupNorm group sums: sum over the group's channels of gn_w * (sum dn, sum dn*xhat)
            float t1 = gw_c * sv[2], t2 = gw_c * sv[3];
            for (int o = cg >> 1; o > 0; o >>= 1) {
                t1 += __shfl_xor_sync(0xffffffffu, t1, o);
                t2 += __shfl_xor_sync(0xffffffffu, t2, o);
            }
            if (tid % cg == 0) {
                s_pub[(tid / cg) * 2 + 0] = t1;
                s_pub[(tid / cg) * 2 + 1] = t2;
            }
        }
        __syncthreads();
        if (tid < 16)
            st_relaxed_u64(xchg_slot(sync, b, j_cta) + tid,
                           ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(s_pub[tid]));

        // ---------------- the group's sums -> d_raw of my rows ----------------
        for (int i = tid; i < G * 16; i += 256) {
            const unsigned long long* src = xchg_slot(sync, b, i >> 4) + (i & 15);
            unsigned long long pk = ld_relaxed_u64(src);
            if ((unsigned int)(pk >> 32) != epoch) {
                const long long t0 = clock64();
                do {
                    pk = ld_relaxed_u64(src);
                    if (clock64() - t0 > 4000000000LL) gbf_timeout(b, i >> 4);
                } while ((unsigned int)(pk >> 32) != epoch);
            }
            s_x[i] = __uint_as_float((unsigned int)pk);
        }
        __syncthreads();
        float m1 = 0.0f, m2 = 0.0f;
        for (int s = 0; s < G; ++s) {
            m1 += s_x[s * 16 + g * 2];
            m2 += s_x[s * 16 + g * 2 + 1];
        }
        m1 *= inv_n;
        m2 *= inv_n;
        const f32x2 nm1 = pkf2(-m1, -m1), nm2 = pkf2(-m2, -m2);
        f32x2 sbs = 0ull;
        for (int rb = tr; rb < R; rb += 2 * n_tr) {
#pragma unroll
            for (int uu = 0; uu < 2; ++uu) {
                const int r = rb + uu * n_tr;
                if (r >= R) break;
                const uint32_t xw = s_raw[r * n_pair + pr];
                f32x2 dv = has_do ? bf2_lo(s_do[r * n_pair + pr]) : 0ull;
                if (has_pool) dv = ffma2(bf2_lo(s_pool[(r >> 1) * n_pair + pr]), half2, dv);
                const f32x2 x = bf2_lo(xw);
                f32x2 z, act, dact;
                sg_silu_pair(x, hA, hB, z, act, dact);
                const f32x2 dn = fmul2(fmul2(dv, Gm), dact);
                const f32x2 xh = ffma2(x, rs2, xo2);
                const f32x2 dz = fmul2(ffma2(xh, nm2, ffma2(dn, GW2, nm1)), rs2);
                sbs = fadd2(sbs, dz);
                float lo, hi;
                upk2(dz, lo, hi);
                s_raw[r * n_pair + pr] = pack_bf16x2(lo, hi);
            }
        }
        {
            float lo, hi;
            upk2(sbs, lo, hi);
            red[(size_t)tr * C + 2 * pr] = lo;
            red[(size_t)tr * C + 2 * pr + 1] = hi;
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            bulk_store(d_raw + ((size_t)b * L + r0) * C, smem_u32(smem), (uint32_t)R * C * 2);
            tma_commit();
        }
        if (tid < C) {
            float s = 0.0f;
            for (int t = 0; t < n_tr; ++t) s += red[(size_t)t * C + tid];
            partial_bias[((size_t)b * G + j_cta) * C + tid] = s;
        }
        __syncthreads();                              // red / s_x / s_pub are reused by the next sample
    }
    if (tid == 0) {
        tma_wait_all<0>();
        xchg_finish(ctrl);
    }
}

static int gbf_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    }
    return n;
}

// CTAs per sample of the one-pass kernel for this layer, 0 if the shape needs the two-pass kernels
int gn_bwd_fused_group(int L, int C, int Cc, bool has_do, bool has_pool) {
    GbfPlan pl;
    if (!g_gn_bwd_fused || !gbf_plan(L, C, Cc, has_do, has_pool, &pl)) return 0;
    return pl.G;
}

// partial [B, G, C, 4+Cc], partial_bias [B, G, C]; GW_ERR_UNSUPPORTED if the group does not fit on the GPU
int gn_bwd_fused(const GnBwdArgs& a, int B, float* partial, float* partial_bias, void* d_raw, void* sync, cudaStream_t st) {
    GbfPlan pl;
    if (!gbf_plan(a.L, a.C, a.Cc, a.do_a != nullptr, a.do_pool != nullptr, &pl)) return GW_ERR_UNSUPPORTED;
    const int Cc = a.Cc;
    // every CTA of a group spins on its peers: the grid must be co-resident -> size it from the occupancy of this variant
#define GBF_GO(CCV)                                                                                                       \
    do {                                                                                                                  \
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_fused_kernel<CCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem)); \
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_fused_kernel<CCV>, cudaFuncAttributePreferredSharedMemoryCarveout,            \
                                     (int)cudaSharedmemCarveoutMaxShared));                                              \
        int occ = 0;                                                                                                      \
        GW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_bwd_fused_kernel<CCV>, 256, pl.smem));             \
        int n_groups = occ * gbf_sm_count() / pl.G;                                                                       \
        if (n_groups > B) n_groups = B;                                                                                   \
        if (n_groups < 1) return GW_ERR_UNSUPPORTED;                                                                      \
        GW_CUDA(gw_launch_pdl(gn_bwd_fused_kernel<CCV>, pl.G * n_groups, dim3(256), (size_t)(pl.smem), st, a, pl, B, n_groups, partial, partial_bias,        \
                                                                         (bf16*)d_raw, sync));                             \
    } while (0)
    if (Cc == 0) GBF_GO(0);
    else if (Cc == 1) GBF_GO(1);
    else if (Cc == 5) GBF_GO(5);
    else GBF_GO(-1);
#undef GBF_GO
    GW_LAUNCH_CHECK();
    return GW_OK;
}
