// HBM-streaming versions of the fused GroupNorm kernels for bf16 channels-last activations (sm_100a).
//
// The register-streaming kernels in forward.cu / backward.cu keep ~50 KB of loads in flight per SM and top out near 70 % of
// the measured copy bandwidth.  Here the data path is decoupled from the register file: a CTA owns a contiguous row range
// of one sample (channels-last => one contiguous byte range), an elected thread streams it through a 4-deep ring of 8 KB
// shared-memory stages with 1-D bulk copies (cp.async.bulk + mbarrier complete_tx), all 256 threads compute from shared
// memory, results are staged in shared memory and leave through bulk stores.  ~100 KB per SM are in flight regardless of
// register pressure.
#include "common.cuh"
#include "../../include/gwb200.h"
#include "tc_common.cuh"

#define SG_STAGE_BYTES 8192          // one stage of raw rows: S rows x C channels x 2 B
#define SG_DEPTH 4
#define SG_MAX_CC 8

// ------------------------------------------------------------------------------------------------
// forward: GroupNorm-apply + SiLU + cond 1x1 conv + FiLM (+ avg_pool), see gn_apply_kernel in forward.cu
// ------------------------------------------------------------------------------------------------
// CV > 0: channel count known at compile time (every stride an immediate, full stages unrolled); CV = 0: generic
template <int CC, int CV>
__global__ void __launch_bounds__(256)
gn_apply_stream_kernel(const bf16* __restrict__ raw, const float* __restrict__ part, int n_part, int L, int C_rt,
                       const float* __restrict__ gn_w, const float* __restrict__ gn_b, const float* __restrict__ cond, int Cc_rt,
                       const float* __restrict__ wc, const float* __restrict__ bc, const float* __restrict__ film, int film_off,
                       long film_b_stride, long film_step_stride, const int* __restrict__ step_ptr, bf16* __restrict__ out,
                       bf16* __restrict__ pooled, float* __restrict__ stats_out, int rows_per_cta) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int NC = CC >= 0 ? CC : SG_MAX_CC;
    constexpr int NCA = NC > 0 ? NC : 1;
    const int Cc = CC >= 0 ? CC : Cc_rt;
    const int C = CV > 0 ? CV : C_rt;
    extern __shared__ __align__(128) uint8_t smem[];
    // layout: in[D][8 KB] | out[2][8 KB] | pool[2][4 KB] | cond[D][S*Cc*4 rounded to 128] | barriers
    const int S = SG_STAGE_BYTES / (C * 2);                       // rows per stage (64 / 32 / 16 for C = 64 / 128 / 256)
    const uint32_t cond_stage = (uint32_t)((S * NCA * 4 + 127) & ~127);
    uint8_t* s_in = smem;
    uint8_t* s_out = s_in + SG_DEPTH * SG_STAGE_BYTES;
    uint8_t* s_pool = s_out + 2 * SG_STAGE_BYTES;
    uint8_t* s_cond = s_pool + 2 * (SG_STAGE_BYTES / 2);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_cond + SG_DEPTH * cond_stage);
    __shared__ float s_mean[8], s_rstd[8];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cg = C / 8;
    const int r0 = blockIdx.x * rows_per_cta;
    const int rows_here = min(rows_per_cta, L - r0);
    const int n_sub = (rows_here + S - 1) / S;
    const bf16* rbase = raw + ((size_t)b * L + r0) * C;
    const float* cbase = cond + ((size_t)b * L + r0) * Cc;

    auto issue_load = [&](int i) {
        const int st = i % SG_DEPTH;
        const int rows_i = min(S, rows_here - i * S);
        const uint32_t bar = smem_u32(bars + st);
        const uint32_t nb = (uint32_t)rows_i * C * 2, ncb = (uint32_t)rows_i * Cc * 4;
        mbar_expect_tx(bar, nb + (NC > 0 ? ncb : 0u));
        bulk_load(smem_u32(s_in + st * SG_STAGE_BYTES), rbase + (size_t)i * S * C, nb, bar);
        if (NC > 0) bulk_load(smem_u32(s_cond + st * cond_stage), cbase + (size_t)i * S * Cc, ncb, bar);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < SG_DEPTH; ++s) mbar_init(smem_u32(bars + s), 1);
        fence_barrier_init();
        fence_proxy_async();
        for (int i = 0; i < SG_DEPTH && i < n_sub; ++i) issue_load(i);      // the stream starts before the statistics are ready
    }
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
    const int n_quad = C / 4;                                         // 16 / 32 / 64: divides 256
    const int quad = threadIdx.x % n_quad, pr0 = threadIdx.x / n_quad, pr_stride = 256 / n_quad;
    const int step = step_ptr != nullptr ? *step_ptr : 0;
    const float* fr = film + (size_t)step * film_step_stride + (size_t)b * film_b_stride + film_off;
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
    const f32x2 half2 = pkf2(0.5f, 0.5f);
    const bool do_pool = pooled != nullptr;
    const int Lp = L / 2;
    bf16* obase = out + ((size_t)b * L + r0) * C;
    bf16* pbase = do_pool ? pooled + ((size_t)b * Lp + (r0 >> 1)) * C : nullptr;

    for (int i = 0; i < n_sub; ++i) {
        const int st = i % SG_DEPTH;
        const int rows_i = min(S, rows_here - i * S);
        mbar_wait(smem_u32(bars + st), (uint32_t)((i / SG_DEPTH) & 1));
        const uint8_t* in = s_in + st * SG_STAGE_BYTES;
        const float* cd = reinterpret_cast<const float*>(s_cond + st * cond_stage);
        uint8_t* so = s_out + (i & 1) * SG_STAGE_BYTES;
        uint8_t* sp = s_pool + (i & 1) * (SG_STAGE_BYTES / 2);
        auto pair_rows = [&](int pr) {
            f32x2 o[2][2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = 2 * pr + h;                         // rows_i is even on this path
                const uint2 xr = *reinterpret_cast<const uint2*>(in + ((size_t)r * C + quad * 4) * 2);
                const uint32_t w2[2] = {xr.x, xr.y};
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
                        for (int j = 0; j < NCA; ++j) {
                            const float cvj = j < Cc ? cd[r * Cc + j] : 0.0f;
                            v = ffma2(W[j][q], pkf2(cvj, cvj), v);
                        }
                    }
                    o[h][q] = v;
                }
                float a0, a1, a2, a3;
                upk2(o[h][0], a0, a1);
                upk2(o[h][1], a2, a3);
                uint2 rr;
                rr.x = pack_bf16x2(a0, a1);
                rr.y = pack_bf16x2(a2, a3);
                *reinterpret_cast<uint2*>(so + ((size_t)r * C + quad * 4) * 2) = rr;
            }
            if (do_pool) {
                float a0, a1, a2, a3;
                upk2(fmul2(fadd2(o[0][0], o[1][0]), half2), a0, a1);
                upk2(fmul2(fadd2(o[0][1], o[1][1]), half2), a2, a3);
                uint2 rr;
                rr.x = pack_bf16x2(a0, a1);
                rr.y = pack_bf16x2(a2, a3);
                *reinterpret_cast<uint2*>(sp + ((size_t)pr * C + quad * 4) * 2) = rr;
            }
        };
        if (CV > 0 && rows_i == S) {
            constexpr int TRIPS = CV > 0 ? (SG_STAGE_BYTES / (CV * 2) / 2) / (256 / (CV / 4)) : 1;
#pragma unroll
            for (int kq = 0; kq < TRIPS; ++kq) pair_rows(pr0 + kq * pr_stride);
        } else {
            for (int pr = pr0; 2 * pr < rows_i; pr += pr_stride) pair_rows(pr);
        }
        fence_proxy_async();                                          // my smem writes -> visible to the bulk-store engine
        if (threadIdx.x == 0) tma_wait_read<0>();                     // stores of iteration i-1 have drained staging[(i+1)&1]
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_store(obase + (size_t)i * S * C, smem_u32(so), (uint32_t)rows_i * C * 2);
            if (do_pool) bulk_store(pbase + (size_t)i * (S / 2) * C, smem_u32(sp), (uint32_t)(rows_i / 2) * C * 2);
            tma_commit();
            if (i + SG_DEPTH < n_sub) issue_load(i + SG_DEPTH);       // every thread is done reading stage st
        }
    }
    if (threadIdx.x == 0) tma_wait_read<0>();
}

static int gn_stream_rows(int L, int C) {
    const int S = SG_STAGE_BYTES / (C * 2);
    int rows = 8 * S;                                   // 64 KB of raw rows per CTA
    if (rows > L) rows = (L + S - 1) / S * S;
    return rows;
}

// same contract as gw_gn_apply for dtype = GW_BF16; needs L % 4 == 0 and C in {64, 128, 256}
extern "C" int gw_gn_apply_stream(const void* raw, const float* part, int n_part, int B, int L, int C, const float* gn_w,
                                  const float* gn_b, const float* cond, int Cc, const float* wc, const float* bc,
                                  const float* film, int film_off, long film_b_stride, long film_step_stride,
                                  const int* step_ptr, void* out, void* pooled, float* stats_out, void* stream) {
    GW_REQUIRE(C == 64 || C == 128 || C == 256, "gw_gn_apply_stream: C=%d", C);
    GW_REQUIRE(L % 4 == 0 && L >= 4, "gw_gn_apply_stream: L=%d must be a multiple of 4", L);
    GW_REQUIRE(Cc >= 0 && Cc <= SG_MAX_CC, "gw_gn_apply_stream: Cc=%d", Cc);
    GW_REQUIRE((cond != nullptr) == (Cc > 0), "gw_gn_apply_stream: cond/Cc mismatch");
    const int rows = gn_stream_rows(L, C);
    const int S = SG_STAGE_BYTES / (C * 2);
    const int nca = Cc > 0 ? (Cc == 1 || Cc == 5 ? Cc : SG_MAX_CC) : 1;
    const int cond_stage = (S * nca * 4 + 127) & ~127;
    const size_t smem = (size_t)SG_DEPTH * SG_STAGE_BYTES + 2 * SG_STAGE_BYTES + SG_STAGE_BYTES + SG_DEPTH * cond_stage + 64;
    dim3 grid(gw_cdiv(L, rows), B);
    cudaStream_t st = (cudaStream_t)stream;
#define SGA_GO(CCV, CV)                                                                                                    \
    do {                                                                                                                   \
        GW_CUDA(cudaFuncSetAttribute(gn_apply_stream_kernel<CCV, CV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        GW_CUDA(cudaFuncSetAttribute(gn_apply_stream_kernel<CCV, CV>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));   \
        GW_CUDA(gw_launch_pdl(gn_apply_stream_kernel<CCV, CV>, grid, dim3(256), (size_t)(smem), st, (const bf16*)raw, part, n_part, L, C, gn_w, gn_b, cond, Cc, wc, bc, \
                                                             film, film_off, film_b_stride, film_step_stride, step_ptr,        \
                                                             (bf16*)out, (bf16*)pooled, stats_out, rows));                      \
    } while (0)
#define SGA_C(CCV)                          \
    do {                                    \
        if (C == 64) SGA_GO(CCV, 64);       \
        else if (C == 128) SGA_GO(CCV, 128); \
        else SGA_GO(CCV, 256);              \
    } while (0)
    if (Cc == 0) SGA_GO(0, 0);
    else if (Cc == 1) SGA_C(1);
    else if (Cc == 5) SGA_C(5);
    else SGA_GO(-1, 0);
#undef SGA_C
#undef SGA_GO
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ================================================================================================
// backward: the two passes of gw_gn_bwd (backward.cu) as streaming kernels
// ================================================================================================
#include "gn_bwd.cuh"

int g_gn_bwd_stats_fast = 1;
bool gn_bwd_stream_fast_ok(const GnBwdArgs& a) {
    if (!g_gn_bwd_stats_fast || !(a.Cc == 1 || a.Cc == 5)) return false;
    if (a.do_eps != nullptr) return a.C == 64 && a.do_a == nullptr && a.do_pool == nullptr;    // head-gradient source (last decoder)
    return a.do_a != nullptr;
}
int gn_bwd_stream_rows(int L, int C) {
    const int S = SG_STAGE_BYTES / (C * 2);
    int rows = 8 * S;
    if (rows > L) rows = L;
    return rows < 1 ? 1 : rows;
}

// Stage layout (bytes): raw [S*C*2] | do_a [S*C*2] | do_pool [S/2*C*2] | cond [S*Cc*4 -> 128]
struct SgBwdLayout {
    uint32_t off_do, off_pool, off_cond, stage;
};
__host__ __device__ inline SgBwdLayout sg_bwd_layout(int C, int Cc, bool has_do, bool has_pool) {
    const int S = SG_STAGE_BYTES / (C * 2);
    SgBwdLayout l;
    uint32_t o = SG_STAGE_BYTES;
    l.off_do = o;
    if (has_do) o += SG_STAGE_BYTES;
    l.off_pool = o;
    if (has_pool) o += SG_STAGE_BYTES / 2;
    l.off_cond = o;
    if (Cc > 0) o += (uint32_t)((S * Cc * 4 + 127) & ~127);
    l.stage = o;
    return l;
}

// common producer: one sub-chunk of rows [i*S, i*S + rows_i) of this CTA's range into stage i % depth
struct SgBwdStream {
    const GnBwdArgs* a;
    int b, r0, rows_here, S, depth;
    SgBwdLayout lay;
    uint8_t* base;
    uint64_t* bars;
    __device__ void issue(int i, bool want_cond) const {
        const int C = a->C, L = a->L, Cc = a->Cc;
        const int st = i % depth;
        const int rows_i = min(S, rows_here - i * S);
        const uint32_t bar = smem_u32(bars + st);
        const uint32_t nb = (uint32_t)rows_i * C * 2;
        uint32_t total = nb;
        if (a->do_a) total += nb;
        if (a->do_pool) total += nb / 2;
        if (want_cond && Cc > 0) total += (uint32_t)rows_i * Cc * 4;
        mbar_expect_tx(bar, total);
        uint8_t* sb = base + (size_t)st * lay.stage;
        const size_t row = (size_t)b * L + r0 + (size_t)i * S;
        bulk_load(smem_u32(sb), (const bf16*)a->raw + row * C, nb, bar);
        if (a->do_a) bulk_load(smem_u32(sb + lay.off_do), (const bf16*)a->do_a + row * C, nb, bar);
        if (a->do_pool)
            bulk_load(smem_u32(sb + lay.off_pool), (const bf16*)a->do_pool + ((size_t)b * (L / 2) + ((r0 + i * S) >> 1)) * C, nb / 2, bar);
        if (want_cond && Cc > 0) bulk_load(smem_u32(sb + lay.off_cond), a->cond + row * Cc, (uint32_t)rows_i * Cc * 4, bar);
    }
};

template <int CC, bool HEAD>
__global__ void __launch_bounds__(256, 3) gn_bwd_stats_stream_kernel(GnBwdArgs a, float* __restrict__ partial, int depth) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int NC = CC >= 0 ? CC : SG_MAX_CC;
    constexpr int NV = 4 + NC;
    extern __shared__ __align__(128) uint8_t smem[];
    const int Cc = CC >= 0 ? CC : a.Cc;
    const int b = blockIdx.y, C = a.C, L = a.L;
    const int S = SG_STAGE_BYTES / (C * 2);
    SgBwdStream ps;
    ps.a = &a; ps.b = b; ps.r0 = blockIdx.x * a.rows_per_cta; ps.rows_here = min(a.rows_per_cta, L - ps.r0); ps.S = S;
    ps.depth = depth; ps.lay = sg_bwd_layout(C, Cc, a.do_a != nullptr, a.do_pool != nullptr); ps.base = smem;
    ps.bars = reinterpret_cast<uint64_t*>(smem + (size_t)depth * ps.lay.stage);
    const int n_sub = (ps.rows_here + S - 1) / S;
    if (threadIdx.x == 0) {
        for (int s = 0; s < depth; ++s) mbar_init(smem_u32(ps.bars + s), 1);
        fence_barrier_init();
        fence_proxy_async();
        for (int i = 0; i < depth && i < n_sub; ++i) ps.issue(i, true);
    }
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
    const float* de = (HEAD && a.do_eps) ? a.do_eps + (size_t)b * L : nullptr;
    f32x2 wk[3][2] = {{0ull, 0ull}, {0ull, 0ull}, {0ull, 0ull}};
    if (HEAD && de) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                wk[k][h] = pkf2(a.do_w[(quad * 4 + 2 * h) * 3 + k], a.do_w[(quad * 4 + 2 * h + 1) * 3 + k]);
    }
    f32x2 acc[2][NV];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[h][v] = 0ull;
    const f32x2 half2 = pkf2(0.5f, 0.5f);
    const bool has_do = a.do_a != nullptr, has_pool = a.do_pool != nullptr;
    __syncthreads();
    for (int i = 0; i < n_sub; ++i) {
        const int st = i % depth;
        const int rows_i = min(S, ps.rows_here - i * S);
        mbar_wait(smem_u32(ps.bars + st), (uint32_t)((i / depth) & 1));
        const uint8_t* sb = smem + (size_t)st * ps.lay.stage;
        const float* cd = reinterpret_cast<const float*>(sb + ps.lay.off_cond);
        for (int rb = tr; rb < rows_i; rb += 2 * n_tr) {
#pragma unroll
          for (int uu = 0; uu < 2; ++uu) {                            // two independent rows in flight per thread (ILP)
            const int r = rb + uu * n_tr;
            if (r >= rows_i) break;
            const uint2 xr = *reinterpret_cast<const uint2*>(sb + ((size_t)r * C + quad * 4) * 2);
            uint2 dr = make_uint2(0u, 0u), pl = make_uint2(0u, 0u);
            if (has_do) dr = *reinterpret_cast<const uint2*>(sb + ps.lay.off_do + ((size_t)r * C + quad * 4) * 2);
            if (has_pool) pl = *reinterpret_cast<const uint2*>(sb + ps.lay.off_pool + ((size_t)(r >> 1) * C + quad * 4) * 2);
            float e3[3] = {0.0f, 0.0f, 0.0f};
            if (HEAD && de) {
                const int l = ps.r0 + i * S + r;
                e3[0] = l + 1 < L ? de[l + 1] : 0.0f;
                e3[1] = de[l];
                e3[2] = l > 0 ? de[l - 1] : 0.0f;
            }
            const uint32_t xw[2] = {xr.x, xr.y}, dw[2] = {dr.x, dr.y}, pw[2] = {pl.x, pl.y};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                f32x2 dv = has_do ? bf2_lo(dw[h]) : 0ull;
                if (has_pool) dv = ffma2(bf2_lo(pw[h]), half2, dv);
                if (HEAD && de) {
                    const f32x2 t = ffma2(wk[1][h], pkf2(e3[1], e3[1]), fmul2(wk[2][h], pkf2(e3[2], e3[2])));
                    dv = fadd2(dv, ffma2(wk[0][h], pkf2(e3[0], e3[0]), t));
                }
                const f32x2 x = bf2_lo(xw[h]);
                f32x2 z, act, dact;
                sg_silu_pair(x, hA[h], hB[h], z, act, dact);
                const f32x2 dn = fmul2(fmul2(dv, G[h]), dact);
                const f32x2 xh = ffma2(x, rs2, xo2);
                acc[h][0] = fadd2(acc[h][0], dv);
                acc[h][1] = ffma2(dv, act, acc[h][1]);
                acc[h][2] = fadd2(acc[h][2], dn);
                acc[h][3] = ffma2(dn, xh, acc[h][3]);
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    const float cvj = j < Cc ? cd[r * Cc + j] : 0.0f;
                    acc[h][4 + j] = ffma2(dv, pkf2(cvj, cvj), acc[h][4 + j]);
                }
            }
          }
        }
        __syncthreads();                                              // every thread is done with stage st
        if (threadIdx.x == 0 && i + depth < n_sub) ps.issue(i + depth, true);
    }
    // reduce the thread rows through shared memory (the stage ring is free now)
    float* red = reinterpret_cast<float*>(smem);
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

// Same sums with every stride a compile-time constant (C, cond channels, which gradients come in): the generic kernel above
// spends ~250 warp instructions per (row, 4 channels) of which ~85 are arithmetic and is ISSUE-bound at 61 % issue utilisation,
// 2.8 TB/s (profiles/r01j_ncu_gn_bwd_stats.md).  Here a full stage is 4 rows per thread at immediate offsets.
// HEAD (last decoder, C = 64): the incoming gradient is the head conv's, dout[l,c] = sum_k do_w[c,k] * do_eps[l-k+1], formed on the fly
// from the CTA's slice of d_eps in shared memory -- the [B, L, 64] d_h tensor is neither written by gw_final_bwd nor read here.
template <int C, int CC, bool POOL, bool HEAD = false>
__global__ void __launch_bounds__(256, 3) gn_bwd_stats_fast_kernel(GnBwdArgs a, float* __restrict__ partial, int depth) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int S = SG_STAGE_BYTES / (C * 2);                       // rows per stage
    constexpr int NQ = C / 4, NTR = 256 / NQ, RPT = S / NTR;          // channel quads, row lanes, rows per thread and stage
    constexpr int NV = 4 + CC;                                        // [sum do, sum do*act, sum dn, sum dn*xhat, cond..]
    constexpr uint32_t OFF_DO = SG_STAGE_BYTES, OFF_POOL = 2 * SG_STAGE_BYTES;
    constexpr uint32_t OFF_COND = HEAD ? SG_STAGE_BYTES : OFF_POOL + (POOL ? SG_STAGE_BYTES / 2 : 0);
    static_assert(RPT * NTR == S && NTR % 2 == 0 && !(HEAD && POOL), "stage geometry");
    extern __shared__ __align__(128) uint8_t smem[];
    const int b = blockIdx.y, L = a.L;
    SgBwdStream ps;
    ps.a = &a; ps.b = b; ps.r0 = blockIdx.x * a.rows_per_cta; ps.rows_here = min(a.rows_per_cta, L - ps.r0); ps.S = S;
    ps.depth = depth; ps.lay = sg_bwd_layout(C, CC, !HEAD, POOL); ps.base = smem;
    ps.bars = reinterpret_cast<uint64_t*>(smem + (size_t)depth * ps.lay.stage);
    float* s_de = reinterpret_cast<float*>(ps.bars + 8);              // HEAD: d_eps[r0 - 1 .. r0 + rows_here]
    const uint32_t stage_bytes = ps.lay.stage;
    const int n_sub = (ps.rows_here + S - 1) / S;
    if (threadIdx.x == 0) {
        for (int s = 0; s < depth; ++s) mbar_init(smem_u32(ps.bars + s), 1);
        fence_barrier_init();
        fence_proxy_async();
        for (int i = 0; i < depth && i < n_sub; ++i) ps.issue(i, true);
    }
    const int quad = threadIdx.x % NQ, tr = threadIdx.x / NQ;
    f32x2 wk[3][2] = {{0ull, 0ull}, {0ull, 0ull}, {0ull, 0ull}};
    if (HEAD) {
        const float* de = a.do_eps + (size_t)b * L;
        for (int i = threadIdx.x; i < ps.rows_here + 2; i += 256) {
            const int l = ps.r0 - 1 + i;
            s_de[i] = (l >= 0 && l < L) ? de[l] : 0.0f;
        }
#pragma unroll
        for (int kk = 0; kk < 3; ++kk)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                wk[kk][h] = pkf2(a.do_w[(quad * 4 + 2 * h) * 3 + kk], a.do_w[(quad * 4 + 2 * h + 1) * 3 + kk]);
    }
    f32x2 hA[2], hB[2], G[2], rs2, xo2;
    {
        constexpr int cg = C / 8;
        const int g = (quad * 4) / cg;
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
    const f32x2 half2 = pkf2(0.5f, 0.5f);
    // one (row, channel quad): x, incoming gradient, cond values -> the 4 + CC running sums
    auto row = [&](uint2 xr, uint2 dr, uint2 pl, const float* cdr, const float* ep) {
        float cv[CC > 0 ? CC : 1];
#pragma unroll
        for (int j = 0; j < CC; ++j) cv[j] = cdr[j];
        const uint32_t xw[2] = {xr.x, xr.y}, dw[2] = {dr.x, dr.y}, pw[2] = {pl.x, pl.y};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            f32x2 dv;
            if (HEAD) {                                              // ep[0..2] = d_eps[l-1], d_eps[l], d_eps[l+1]
                const float em = ep[0], ec = ep[1], en = ep[2];
                dv = ffma2(wk[0][h], pkf2(en, en), ffma2(wk[1][h], pkf2(ec, ec), fmul2(wk[2][h], pkf2(em, em))));
            } else {
                dv = bf2_lo(dw[h]);
            }
            if (POOL) dv = ffma2(bf2_lo(pw[h]), half2, dv);
            const f32x2 x = bf2_lo(xw[h]);
            f32x2 z, act, dact;
            sg_silu_pair(x, hA[h], hB[h], z, act, dact);
            const f32x2 dn = fmul2(fmul2(dv, G[h]), dact);
            const f32x2 xh = ffma2(x, rs2, xo2);
            acc[h][0] = fadd2(acc[h][0], dv);
            acc[h][1] = ffma2(dv, act, acc[h][1]);
            acc[h][2] = fadd2(acc[h][2], dn);
            acc[h][3] = ffma2(dn, xh, acc[h][3]);
#pragma unroll
            for (int j = 0; j < CC; ++j) acc[h][4 + j] = ffma2(dv, pkf2(cv[j], cv[j]), acc[h][4 + j]);
        }
    };
    const uint32_t t_off = (uint32_t)(tr * C + quad * 4) * 2;          // my quad in row tr of a stage
    const uint32_t p_off = OFF_POOL + (uint32_t)((tr >> 1) * C + quad * 4) * 2;
    __syncthreads();
    for (int i = 0; i < n_sub; ++i) {
        const int st = i % depth;
        const int rows_i = min(S, ps.rows_here - i * S);
        mbar_wait(smem_u32(ps.bars + st), (uint32_t)((i / depth) & 1));
        const uint8_t* sb = smem + (size_t)st * stage_bytes;
        const float* cd = reinterpret_cast<const float*>(sb + OFF_COND) + tr * CC;
        if (rows_i == S) {
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const uint2 xr = *reinterpret_cast<const uint2*>(sb + t_off + k * NTR * C * 2);
                uint2 dr = make_uint2(0u, 0u);
                if (!HEAD) dr = *reinterpret_cast<const uint2*>(sb + OFF_DO + t_off + k * NTR * C * 2);
                uint2 pl = make_uint2(0u, 0u);
                if (POOL) pl = *reinterpret_cast<const uint2*>(sb + p_off + k * (NTR / 2) * C * 2);
                row(xr, dr, pl, cd + k * NTR * CC, s_de + i * S + tr + k * NTR);
            }
        } else {
            for (int r = tr; r < rows_i; r += NTR) {
                const uint2 xr = *reinterpret_cast<const uint2*>(sb + ((size_t)r * C + quad * 4) * 2);
                uint2 dr = make_uint2(0u, 0u);
                if (!HEAD) dr = *reinterpret_cast<const uint2*>(sb + OFF_DO + ((size_t)r * C + quad * 4) * 2);
                uint2 pl = make_uint2(0u, 0u);
                if (POOL) pl = *reinterpret_cast<const uint2*>(sb + OFF_POOL + ((size_t)(r >> 1) * C + quad * 4) * 2);
                row(xr, dr, pl, reinterpret_cast<const float*>(sb + OFF_COND) + r * CC, s_de + i * S + r);
            }
        }
        __syncthreads();                                              // every thread is done with stage st
        if (threadIdx.x == 0 && i + depth < n_sub) ps.issue(i + depth, true);
    }
    // reduce the thread rows through shared memory (the stage ring is free now)
    float* red = reinterpret_cast<float*>(smem);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            float lo, hi;
            upk2(acc[h][v], lo, hi);
            red[((size_t)tr * C + quad * 4 + 2 * h) * NV + v] = lo;
            red[((size_t)tr * C + quad * 4 + 2 * h + 1) * NV + v] = hi;
        }
    __syncthreads();
    float* pt = partial + ((size_t)b * gridDim.x + blockIdx.x) * C * NV;
    for (int i = threadIdx.x; i < C * NV; i += 256) {
        float sacc = 0.0f;
#pragma unroll
        for (int t = 0; t < NTR; ++t) sacc += red[(size_t)t * C * NV + i];
        pt[i] = sacc;
    }
}

template <bool HEAD>
__global__ void __launch_bounds__(256, 3) gn_bwd_apply_stream_kernel(GnBwdArgs a, const float* __restrict__ gstat,
                                                                  bf16* __restrict__ d_raw, float* __restrict__ partial_bias,
                                                                  int depth) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(128) uint8_t smem[];
    const int b = blockIdx.y, C = a.C, L = a.L;
    const int S = SG_STAGE_BYTES / (C * 2);
    SgBwdStream ps;
    ps.a = &a; ps.b = b; ps.r0 = blockIdx.x * a.rows_per_cta; ps.rows_here = min(a.rows_per_cta, L - ps.r0); ps.S = S;
    ps.depth = depth; ps.lay = sg_bwd_layout(C, 0, a.do_a != nullptr, a.do_pool != nullptr); ps.base = smem;
    uint8_t* s_out = smem + (size_t)depth * ps.lay.stage;             // [2][8 KB] output staging
    ps.bars = reinterpret_cast<uint64_t*>(s_out + 2 * SG_STAGE_BYTES);
    const int n_sub = (ps.rows_here + S - 1) / S;
    if (threadIdx.x == 0) {
        for (int s = 0; s < depth; ++s) mbar_init(smem_u32(ps.bars + s), 1);
        fence_barrier_init();
        fence_proxy_async();
        for (int i = 0; i < depth && i < n_sub; ++i) ps.issue(i, false);
    }
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
    const float* de = (HEAD && a.do_eps) ? a.do_eps + (size_t)b * L : nullptr;
    f32x2 wk[3][2] = {{0ull, 0ull}, {0ull, 0ull}, {0ull, 0ull}};
    if (HEAD && de) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                wk[k][h] = pkf2(a.do_w[(quad * 4 + 2 * h) * 3 + k], a.do_w[(quad * 4 + 2 * h + 1) * 3 + k]);
    }
    f32x2 sbs[2] = {0ull, 0ull};
    const f32x2 half2 = pkf2(0.5f, 0.5f);
    const bool has_do = a.do_a != nullptr, has_pool = a.do_pool != nullptr;
    bf16* obase = d_raw + ((size_t)b * L + ps.r0) * C;
    __syncthreads();
    for (int i = 0; i < n_sub; ++i) {
        const int st = i % depth;
        const int rows_i = min(S, ps.rows_here - i * S);
        mbar_wait(smem_u32(ps.bars + st), (uint32_t)((i / depth) & 1));
        const uint8_t* sb = smem + (size_t)st * ps.lay.stage;
        uint8_t* so = s_out + (i & 1) * SG_STAGE_BYTES;
        for (int rb = tr; rb < rows_i; rb += 2 * n_tr) {
#pragma unroll
          for (int uu = 0; uu < 2; ++uu) {                            // two independent rows in flight per thread (ILP)
            const int r = rb + uu * n_tr;
            if (r >= rows_i) break;
            const uint2 xr = *reinterpret_cast<const uint2*>(sb + ((size_t)r * C + quad * 4) * 2);
            uint2 dr = make_uint2(0u, 0u), pl = make_uint2(0u, 0u);
            if (has_do) dr = *reinterpret_cast<const uint2*>(sb + ps.lay.off_do + ((size_t)r * C + quad * 4) * 2);
            if (has_pool) pl = *reinterpret_cast<const uint2*>(sb + ps.lay.off_pool + ((size_t)(r >> 1) * C + quad * 4) * 2);
            float e3[3] = {0.0f, 0.0f, 0.0f};
            if (HEAD && de) {
                const int l = ps.r0 + i * S + r;
                e3[0] = l + 1 < L ? de[l + 1] : 0.0f;
                e3[1] = de[l];
                e3[2] = l > 0 ? de[l - 1] : 0.0f;
            }
            const uint32_t xw[2] = {xr.x, xr.y}, dw[2] = {dr.x, dr.y}, pw[2] = {pl.x, pl.y};
            uint2 outw;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                f32x2 dv = has_do ? bf2_lo(dw[h]) : 0ull;
                if (has_pool) dv = ffma2(bf2_lo(pw[h]), half2, dv);
                if (HEAD && de) {
                    const f32x2 t = ffma2(wk[1][h], pkf2(e3[1], e3[1]), fmul2(wk[2][h], pkf2(e3[2], e3[2])));
                    dv = fadd2(dv, ffma2(wk[0][h], pkf2(e3[0], e3[0]), t));
                }
                const f32x2 x = bf2_lo(xw[h]);
                f32x2 z, act, dact;
                sg_silu_pair(x, hA[h], hB[h], z, act, dact);
                const f32x2 dn = fmul2(fmul2(dv, G[h]), dact);
                const f32x2 xh = ffma2(x, rs2, xo2);
                const f32x2 dz = fmul2(ffma2(xh, nm2, ffma2(dn, GW2[h], nm1)), rs2);
                sbs[h] = fadd2(sbs[h], dz);
                float lo, hi;
                upk2(dz, lo, hi);
                if (h == 0) outw.x = pack_bf16x2(lo, hi); else outw.y = pack_bf16x2(lo, hi);
            }
            *reinterpret_cast<uint2*>(so + ((size_t)r * C + quad * 4) * 2) = outw;
          }
        }
        fence_proxy_async();
        if (threadIdx.x == 0) tma_wait_read<0>();
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_store(obase + (size_t)i * S * C, smem_u32(so), (uint32_t)rows_i * C * 2);
            tma_commit();
            if (i + depth < n_sub) ps.issue(i + depth, false);
        }
    }
    if (threadIdx.x == 0) tma_wait_read<0>();
    __syncthreads();
    float* red = reinterpret_cast<float*>(smem);                      // stage ring is free (loads done, stores read staging only)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float lo, hi;
        upk2(sbs[h], lo, hi);
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

// apply pass with compile-time strides (see gn_bwd_stats_fast_kernel)
template <int C, bool POOL, bool HEAD = false>
__global__ void __launch_bounds__(256, 3) gn_bwd_apply_fast_kernel(GnBwdArgs a, const float* __restrict__ gstat,
                                                                bf16* __restrict__ d_raw, float* __restrict__ partial_bias,
                                                                int depth) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int S = SG_STAGE_BYTES / (C * 2);
    constexpr int NQ = C / 4, NTR = 256 / NQ, RPT = S / NTR;
    constexpr uint32_t OFF_DO = SG_STAGE_BYTES, OFF_POOL = 2 * SG_STAGE_BYTES;
    extern __shared__ __align__(128) uint8_t smem[];
    const int b = blockIdx.y, L = a.L;
    SgBwdStream ps;
    ps.a = &a; ps.b = b; ps.r0 = blockIdx.x * a.rows_per_cta; ps.rows_here = min(a.rows_per_cta, L - ps.r0); ps.S = S;
    ps.depth = depth; ps.lay = sg_bwd_layout(C, 0, !HEAD, POOL); ps.base = smem;
    const uint32_t stage_bytes = ps.lay.stage;
    uint8_t* s_out = smem + (size_t)depth * stage_bytes;              // [2][8 KB] output staging
    ps.bars = reinterpret_cast<uint64_t*>(s_out + 2 * SG_STAGE_BYTES);
    float* s_de = reinterpret_cast<float*>(ps.bars + 8);              // HEAD: d_eps[r0 - 1 .. r0 + rows_here]
    const int n_sub = (ps.rows_here + S - 1) / S;
    if (threadIdx.x == 0) {
        for (int s = 0; s < depth; ++s) mbar_init(smem_u32(ps.bars + s), 1);
        fence_barrier_init();
        fence_proxy_async();
        for (int i = 0; i < depth && i < n_sub; ++i) ps.issue(i, false);
    }
    const int quad = threadIdx.x % NQ, tr = threadIdx.x / NQ;
    f32x2 wk[3][2] = {{0ull, 0ull}, {0ull, 0ull}, {0ull, 0ull}};
    if (HEAD) {
        const float* de = a.do_eps + (size_t)b * L;
        for (int i = threadIdx.x; i < ps.rows_here + 2; i += 256) {
            const int l = ps.r0 - 1 + i;
            s_de[i] = (l >= 0 && l < L) ? de[l] : 0.0f;
        }
#pragma unroll
        for (int kk = 0; kk < 3; ++kk)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                wk[kk][h] = pkf2(a.do_w[(quad * 4 + 2 * h) * 3 + kk], a.do_w[(quad * 4 + 2 * h + 1) * 3 + kk]);
    }
    f32x2 hA[2], hB[2], G[2], GW2[2], rs2, xo2, nm1, nm2;
    {
        constexpr int cg = C / 8;
        const int g = (quad * 4) / cg;
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
    f32x2 sbs[2] = {0ull, 0ull};
    const f32x2 half2 = pkf2(0.5f, 0.5f);
    auto row = [&](uint2 xr, uint2 dr, uint2 pl, const float* ep) -> uint2 {
        const uint32_t xw[2] = {xr.x, xr.y}, dw[2] = {dr.x, dr.y}, pw[2] = {pl.x, pl.y};
        uint32_t ow[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            f32x2 dv;
            if (HEAD) {
                const float em = ep[0], ec = ep[1], en = ep[2];
                dv = ffma2(wk[0][h], pkf2(en, en), ffma2(wk[1][h], pkf2(ec, ec), fmul2(wk[2][h], pkf2(em, em))));
            } else {
                dv = bf2_lo(dw[h]);
            }
            if (POOL) dv = ffma2(bf2_lo(pw[h]), half2, dv);
            const f32x2 x = bf2_lo(xw[h]);
            f32x2 z, act, dact;
            sg_silu_pair(x, hA[h], hB[h], z, act, dact);
            const f32x2 dn = fmul2(fmul2(dv, G[h]), dact);
            const f32x2 xh = ffma2(x, rs2, xo2);
            const f32x2 dz = fmul2(ffma2(xh, nm2, ffma2(dn, GW2[h], nm1)), rs2);
            sbs[h] = fadd2(sbs[h], dz);
            float lo, hi;
            upk2(dz, lo, hi);
            ow[h] = pack_bf16x2(lo, hi);
        }
        return make_uint2(ow[0], ow[1]);
    };
    const uint32_t t_off = (uint32_t)(tr * C + quad * 4) * 2;
    const uint32_t p_off = OFF_POOL + (uint32_t)((tr >> 1) * C + quad * 4) * 2;
    bf16* obase = d_raw + ((size_t)b * L + ps.r0) * C;
    __syncthreads();
    for (int i = 0; i < n_sub; ++i) {
        const int st = i % depth;
        const int rows_i = min(S, ps.rows_here - i * S);
        mbar_wait(smem_u32(ps.bars + st), (uint32_t)((i / depth) & 1));
        const uint8_t* sb = smem + (size_t)st * stage_bytes;
        uint8_t* so = s_out + (i & 1) * SG_STAGE_BYTES;
        if (rows_i == S) {
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const uint2 xr = *reinterpret_cast<const uint2*>(sb + t_off + k * NTR * C * 2);
                uint2 dr = make_uint2(0u, 0u);
                if (!HEAD) dr = *reinterpret_cast<const uint2*>(sb + OFF_DO + t_off + k * NTR * C * 2);
                uint2 pl = make_uint2(0u, 0u);
                if (POOL) pl = *reinterpret_cast<const uint2*>(sb + p_off + k * (NTR / 2) * C * 2);
                *reinterpret_cast<uint2*>(so + t_off + k * NTR * C * 2) = row(xr, dr, pl, s_de + i * S + tr + k * NTR);
            }
        } else {
            for (int r = tr; r < rows_i; r += NTR) {
                const uint2 xr = *reinterpret_cast<const uint2*>(sb + ((size_t)r * C + quad * 4) * 2);
                uint2 dr = make_uint2(0u, 0u);
                if (!HEAD) dr = *reinterpret_cast<const uint2*>(sb + OFF_DO + ((size_t)r * C + quad * 4) * 2);
                uint2 pl = make_uint2(0u, 0u);
                if (POOL) pl = *reinterpret_cast<const uint2*>(sb + OFF_POOL + ((size_t)(r >> 1) * C + quad * 4) * 2);
                *reinterpret_cast<uint2*>(so + ((size_t)r * C + quad * 4) * 2) = row(xr, dr, pl, s_de + i * S + r);
            }
        }
        fence_proxy_async();
        if (threadIdx.x == 0) tma_wait_read<0>();
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_store(obase + (size_t)i * S * C, smem_u32(so), (uint32_t)rows_i * C * 2);
            tma_commit();
            if (i + depth < n_sub) ps.issue(i + depth, false);
        }
    }
    if (threadIdx.x == 0) tma_wait_read<0>();
    if (partial_bias == nullptr) return;                              // conv-bias gradient formed analytically (gn_bwd_param_kernel)
    __syncthreads();
    float* red = reinterpret_cast<float*>(smem);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float lo, hi;
        upk2(sbs[h], lo, hi);
        red[(size_t)tr * C + quad * 4 + 2 * h] = lo;
        red[(size_t)tr * C + quad * 4 + 2 * h + 1] = hi;
    }
    __syncthreads();
    float* pt = partial_bias + ((size_t)b * gridDim.x + blockIdx.x) * C;
    for (int i = threadIdx.x; i < C; i += 256) {
        float sacc = 0.0f;
#pragma unroll
        for (int t = 0; t < NTR; ++t) sacc += red[(size_t)t * C + i];
        pt[i] = sacc;
    }
}

int gn_bwd_stats_stream(const GnBwdArgs& a, int B, float* partial, cudaStream_t st) {
    const int C = a.C, Cc = a.Cc;
    const SgBwdLayout lay = sg_bwd_layout(C, Cc, a.do_a != nullptr, a.do_pool != nullptr);
    const int n_tr = 256 / (C / 4), nvr = 4 + Cc;
    const int depth = lay.stage > 20000 ? 3 : 4;        // <= 75 KB per CTA: three CTAs per SM
    size_t smem = (size_t)depth * lay.stage + 64;
    const size_t red_bytes = (size_t)n_tr * C * nvr * sizeof(float);
    if (smem < red_bytes) smem = red_bytes;
    dim3 grid(gw_cdiv(a.L, a.rows_per_cta), B);
    const bool head = a.do_eps != nullptr;
    if (gn_bwd_stream_fast_ok(a)) {
        if (head) smem += (size_t)(a.rows_per_cta + 2) * sizeof(float) + 64;      // the CTA's slice of d_eps
        const bool pool = a.do_pool != nullptr;
        if (head) {                                          // last decoder: C = 64, no pooled gradient (gn_bwd_stream_fast_ok)
#define SGH_GO(CCV)                                                                                                         \
    do {                                                                                                                    \
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_stats_fast_kernel<64, CCV, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_stats_fast_kernel<64, CCV, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared)); \
        GW_CUDA(gw_launch_pdl(gn_bwd_stats_fast_kernel<64, CCV, false, true>, grid, dim3(256), (size_t)(smem), st, a, partial, depth));                          \
    } while (0)
            if (Cc == 1) SGH_GO(1); else SGH_GO(5);
#undef SGH_GO
            GW_LAUNCH_CHECK();
            return GW_OK;
        }
#define SGF_GO(CV, CCV, PL)                                                                                                 \
    do {                                                                                                                    \
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_stats_fast_kernel<CV, CCV, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_stats_fast_kernel<CV, CCV, PL>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared)); \
        GW_CUDA(gw_launch_pdl(gn_bwd_stats_fast_kernel<CV, CCV, PL>, grid, dim3(256), (size_t)(smem), st, a, partial, depth));                                   \
    } while (0)
#define SGF_C(CCV, PL)                        \
    do {                                      \
        if (C == 64) SGF_GO(64, CCV, PL);     \
        else if (C == 128) SGF_GO(128, CCV, PL); \
        else SGF_GO(256, CCV, PL);            \
    } while (0)
        if (Cc == 1) { if (pool) SGF_C(1, true); else SGF_C(1, false); }
        else { if (pool) SGF_C(5, true); else SGF_C(5, false); }
#undef SGF_C
#undef SGF_GO
        GW_LAUNCH_CHECK();
        return GW_OK;
    }
#define SGS_GO(CCV, HD)                                                                                                     \
    do {                                                                                                                    \
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_stats_stream_kernel<CCV, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_stats_stream_kernel<CCV, HD>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared)); \
        GW_CUDA(gw_launch_pdl(gn_bwd_stats_stream_kernel<CCV, HD>, grid, dim3(256), (size_t)(smem), st, a, partial, depth));                                     \
    } while (0)
#define SGS_CC(HD)                      \
    do {                                \
        if (Cc == 0) SGS_GO(0, HD);     \
        else if (Cc == 1) SGS_GO(1, HD); \
        else if (Cc == 5) SGS_GO(5, HD); \
        else SGS_GO(-1, HD);            \
    } while (0)
    if (head) SGS_CC(true); else SGS_CC(false);
#undef SGS_CC
#undef SGS_GO
    GW_LAUNCH_CHECK();
    return GW_OK;
}

int gn_bwd_apply_stream(const GnBwdArgs& a, int B, const float* gstat, void* d_raw, float* partial_bias, cudaStream_t st) {
    const int C = a.C;
    const SgBwdLayout lay = sg_bwd_layout(C, 0, a.do_a != nullptr, a.do_pool != nullptr);
    const int depth = lay.stage > 20000 ? 2 : 3;        // + 16 KB of output staging: <= 75 KB per CTA, three CTAs per SM
    size_t smem = (size_t)depth * lay.stage + 2 * SG_STAGE_BYTES + 64;
    dim3 grid(gw_cdiv(a.L, a.rows_per_cta), B);
    if (a.do_eps != nullptr && gn_bwd_stream_fast_ok(a)) {   // head-gradient source (last decoder, C = 64)
        smem += (size_t)(a.rows_per_cta + 2) * sizeof(float) + 64;
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_apply_fast_kernel<64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_apply_fast_kernel<64, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        GW_CUDA(gw_launch_pdl(gn_bwd_apply_fast_kernel<64, false, true>, grid, dim3(256), (size_t)(smem), st, a, gstat, (bf16*)d_raw, partial_bias, depth));
        GW_LAUNCH_CHECK();
        return GW_OK;
    }
    if ((g_gn_bwd_stats_fast && a.do_eps == nullptr && a.do_a != nullptr) || partial_bias == nullptr) {
        const bool pool = a.do_pool != nullptr;
#define SGA_GO(CV, PL)                                                                                                      \
    do {                                                                                                                    \
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_apply_fast_kernel<CV, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_apply_fast_kernel<CV, PL>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared)); \
        GW_CUDA(gw_launch_pdl(gn_bwd_apply_fast_kernel<CV, PL>, grid, dim3(256), (size_t)(smem), st, a, gstat, (bf16*)d_raw, partial_bias, depth));              \
    } while (0)
        if (C == 64) { if (pool) SGA_GO(64, true); else SGA_GO(64, false); }
        else if (C == 128) { if (pool) SGA_GO(128, true); else SGA_GO(128, false); }
        else { if (pool) SGA_GO(256, true); else SGA_GO(256, false); }
#undef SGA_GO
        GW_LAUNCH_CHECK();
        return GW_OK;
    }
    if (a.do_eps != nullptr) {
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_apply_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_apply_stream_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        GW_CUDA(gw_launch_pdl(gn_bwd_apply_stream_kernel<true>, grid, dim3(256), (size_t)(smem), st, a, gstat, (bf16*)d_raw, partial_bias, depth));
    } else {
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_apply_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GW_CUDA(cudaFuncSetAttribute(gn_bwd_apply_stream_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        GW_CUDA(gw_launch_pdl(gn_bwd_apply_stream_kernel<false>, grid, dim3(256), (size_t)(smem), st, a, gstat, (bf16*)d_raw, partial_bias, depth));
    }
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ================================================================================================
// head conv backward (models.py:230) as a streaming kernel for bf16, C = 64 (math: final_bwd_kernel in backward.cu):
//   d_h[l,c] = sum_k wf[c,k] d_eps[l-k+1]   (written when d_h != NULL),   d wf[c,k] += sum_l h[l,c] d_eps[l-k+1],
//   d wf[C,k] += sum_l x_t[l] d_eps[l-k+1],  d bias += sum_l d_eps[l].
// h rows go through the bulk-copy ring, d_h rows leave through staged bulk stores, the CTA's slice of d_eps and x_t sits in shared
// memory.  The register-streaming kernel ran at 2 TB/s (113 us for 233 MB at B = 256, L = 4096).
// partial layout per CTA: [(C+1)*3 + 1] as final_bwd_kernel.
// ================================================================================================
template <bool WRITE_DH>
__global__ void __launch_bounds__(256, 3) final_bwd_stream_kernel(const float* __restrict__ d_eps, const bf16* __restrict__ h,
                                                               const float* __restrict__ net, int Cx, int L,
                                                               const float* __restrict__ wf, bf16* __restrict__ d_h,
                                                               float* __restrict__ partial, int rows_per_cta) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int C = 64, S = SG_STAGE_BYTES / (C * 2), D = 3, NQ = 16, NTR = 16, RPT = S / NTR, NV = C * 3 + 4;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem;                                             // [D][8 KB]
    uint8_t* s_out = smem + D * SG_STAGE_BYTES;                       // [2][8 KB] output staging
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + 2 * SG_STAGE_BYTES);
    float* s_de = reinterpret_cast<float*>(bars + 8);                 // d_eps[r0 - 1 .. r0 + rows_here]
    float* s_xt = s_de + rows_per_cta + 8;                            // x_t[r0 .. r0 + rows_here)
    const int b = blockIdx.y, r0 = blockIdx.x * rows_per_cta;
    const int rows_here = min(rows_per_cta, L - r0);
    const int n_sub = (rows_here + S - 1) / S;
    const bf16* hbase = h + ((size_t)b * L + r0) * C;
    auto issue = [&](int i) {
        const int rows_i = min(S, rows_here - i * S);
        const uint32_t bar = smem_u32(bars + (i % D));
        mbar_expect_tx(bar, (uint32_t)rows_i * C * 2);
        bulk_load(smem_u32(ring + (i % D) * SG_STAGE_BYTES), hbase + (size_t)i * S * C, (uint32_t)rows_i * C * 2, bar);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < D; ++s) mbar_init(smem_u32(bars + s), 1);
        fence_barrier_init();
        fence_proxy_async();
        for (int i = 0; i < D && i < n_sub; ++i) issue(i);
    }
    const float* de = d_eps + (size_t)b * L;
    const float* xr = net + (size_t)b * Cx * L;
    for (int i = threadIdx.x; i < rows_here + 2; i += 256) {
        const int l = r0 - 1 + i;
        s_de[i] = (l >= 0 && l < L) ? de[l] : 0.0f;
    }
    for (int i = threadIdx.x; i < rows_here; i += 256) s_xt[i] = xr[r0 + i];
    const int quad = threadIdx.x % NQ, tr = threadIdx.x / NQ;
    float w[4][3], dw[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            w[i][k] = WRITE_DH ? wf[(quad * 4 + i) * 3 + k] : 0.0f;
            dw[i][k] = 0.0f;
        }
    float ax[4] = {0.0f, 0.0f, 0.0f, 0.0f};                            // quad 0: d wf[C][0..2] (x_t channel), d bias
    bf16* obase = WRITE_DH ? d_h + ((size_t)b * L + r0) * C : nullptr;
    __syncthreads();
    auto row = [&](int rl, uint2 hv) -> uint2 {                       // rl: row inside the CTA's range
        const float em = s_de[rl], ec = s_de[rl + 1], ep = s_de[rl + 2];      // d_eps[l-1], d_eps[l], d_eps[l+1]
        const float hf[4] = {__uint_as_float(hv.x << 16), __uint_as_float(hv.x & 0xffff0000u),
                             __uint_as_float(hv.y << 16), __uint_as_float(hv.y & 0xffff0000u)};
        float o[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (WRITE_DH) o[i] = fmaf(w[i][0], ep, fmaf(w[i][1], ec, w[i][2] * em));
            dw[i][0] = fmaf(hf[i], ep, dw[i][0]);
            dw[i][1] = fmaf(hf[i], ec, dw[i][1]);
            dw[i][2] = fmaf(hf[i], em, dw[i][2]);
        }
        if (quad == 0) {
            const float xv = s_xt[rl];
            ax[0] = fmaf(xv, ep, ax[0]);
            ax[1] = fmaf(xv, ec, ax[1]);
            ax[2] = fmaf(xv, em, ax[2]);
            ax[3] += ec;
        }
        return make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
    };
    const uint32_t t_off = (uint32_t)(tr * C + quad * 4) * 2;
    for (int i = 0; i < n_sub; ++i) {
        const int st = i % D;
        const int rows_i = min(S, rows_here - i * S);
        mbar_wait(smem_u32(bars + st), (uint32_t)((i / D) & 1));
        const uint8_t* sb = ring + st * SG_STAGE_BYTES;
        uint8_t* so = s_out + (i & 1) * SG_STAGE_BYTES;
        if (rows_i == S) {
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const uint2 hv = *reinterpret_cast<const uint2*>(sb + t_off + k * NTR * C * 2);
                const uint2 ov = row(i * S + tr + k * NTR, hv);
                if (WRITE_DH) *reinterpret_cast<uint2*>(so + t_off + k * NTR * C * 2) = ov;
            }
        } else {
            for (int r = tr; r < rows_i; r += NTR) {
                const uint2 hv = *reinterpret_cast<const uint2*>(sb + ((size_t)r * C + quad * 4) * 2);
                const uint2 ov = row(i * S + r, hv);
                if (WRITE_DH) *reinterpret_cast<uint2*>(so + ((size_t)r * C + quad * 4) * 2) = ov;
            }
        }
        if (WRITE_DH) {
            fence_proxy_async();
            if (threadIdx.x == 0) tma_wait_read<0>();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (WRITE_DH) {
                bulk_store(obase + (size_t)i * S * C, smem_u32(so), (uint32_t)rows_i * C * 2);
                tma_commit();
            }
            if (i + D < n_sub) issue(i + D);
        }
    }
    if (threadIdx.x == 0) tma_wait_read<0>();
    __syncthreads();
    float* red = reinterpret_cast<float*>(smem);                      // [NTR][NV] (the ring is free; staging is only read by stores)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) red[(size_t)tr * NV + (quad * 4 + i) * 3 + k] = dw[i][k];
    if (quad == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) red[(size_t)tr * NV + C * 3 + k] = ax[k];
    }
    __syncthreads();
    float* pt = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * NV;
    for (int i = threadIdx.x; i < NV; i += 256) {
        float a = 0.0f;
#pragma unroll
        for (int t = 0; t < NTR; ++t) a += red[(size_t)t * NV + i];
        pt[i] = a;
    }
}

// launcher for gw_final_bwd (backward.cu): bf16, C = 64, L % 4 == 0; *n_cta = partial rows written
int final_bwd_stream(const float* d_eps, const void* h, const float* net, int B, int Cx, int L, const float* wf, void* d_h,
                     float* partial, int* n_cta, cudaStream_t st) {
    const int rows = 512;
    const size_t smem = (size_t)5 * SG_STAGE_BYTES + 64 + (size_t)(2 * rows + 16) * sizeof(float);
    dim3 grid(gw_cdiv(L, rows), B);
    if (d_h != nullptr) {
        GW_CUDA(cudaFuncSetAttribute(final_bwd_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GW_CUDA(cudaFuncSetAttribute(final_bwd_stream_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        GW_CUDA(gw_launch_pdl(final_bwd_stream_kernel<true>, grid, dim3(256), (size_t)(smem), st, d_eps, (const bf16*)h, net, Cx, L, wf, (bf16*)d_h, partial, rows));
    } else {
        GW_CUDA(cudaFuncSetAttribute(final_bwd_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GW_CUDA(cudaFuncSetAttribute(final_bwd_stream_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        GW_CUDA(gw_launch_pdl(final_bwd_stream_kernel<false>, grid, dim3(256), (size_t)(smem), st, d_eps, (const bf16*)h, net, Cx, L, wf, (bf16*)d_h, partial, rows));
    }
    GW_LAUNCH_CHECK();
    *n_cta = grid.x * grid.y;
    return GW_OK;
}

// ================================================================================================
// wgrad of the first conv as a streaming kernel (see wgrad_in_kernel in backward.cu for the math):
//   dW[co][ci][k] = sum_{b,l} d_raw[b,l,co] * x[b,ci,l+k-1]
// d_raw rows go through the bulk-copy ring; the fp32 input rows of the CTA's range sit in shared memory for its lifetime.
// ================================================================================================
template <int CXM>
__global__ void __launch_bounds__(256) wgrad_in_stream_kernel(const float* __restrict__ x, int Cx, int L,
                                                              const bf16* __restrict__ d_raw, float* __restrict__ partial,
                                                              int rows_per_cta) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int C = 64, S = SG_STAGE_BYTES / (C * 2), D = SG_DEPTH;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem;                                         // [D][8 KB]; reused for the final reduction
    const int pitch = rows_per_cta + 8;
    const int nv = Cx * 3;
    const int ring_bytes = max(D * SG_STAGE_BYTES, 8 * C * nv * 4);
    float* xs = reinterpret_cast<float*>(smem + ring_bytes);      // [Cx][pitch]: xs[ci][j] = x[ci][r0 - 1 + j]
    uint64_t* bars = reinterpret_cast<uint64_t*>(xs + Cx * pitch);
    const int b = blockIdx.y, r0 = blockIdx.x * rows_per_cta;
    const int rows_here = min(rows_per_cta, L - r0);
    const int n_sub = (rows_here + S - 1) / S;
    const bf16* dbase = d_raw + ((size_t)b * L + r0) * C;
    auto issue = [&](int i) {
        const int st = i % D;
        const int rows_i = min(S, rows_here - i * S);
        const uint32_t bar = smem_u32(bars + st);
        mbar_expect_tx(bar, (uint32_t)rows_i * C * 2);
        bulk_load(smem_u32(ring + st * SG_STAGE_BYTES), dbase + (size_t)i * S * C, (uint32_t)rows_i * C * 2, bar);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < D; ++s) mbar_init(smem_u32(bars + s), 1);
        fence_barrier_init();
        fence_proxy_async();
        for (int i = 0; i < D && i < n_sub; ++i) issue(i);
    }
    for (int i = threadIdx.x; i < Cx * pitch; i += 256) {
        const int c = i / pitch, p = i % pitch;
        const int l = r0 + p - 1;
        xs[i] = (l >= 0 && l < L) ? x[((size_t)b * Cx + c) * L + l] : 0.0f;
    }
    __syncthreads();
    const int cp = threadIdx.x & 31, tr = threadIdx.x >> 5;      // channel pair, thread row (8)
    float acc[2][CXM * 3];
#pragma unroll
    for (int i = 0; i < CXM * 3; ++i) { acc[0][i] = 0.0f; acc[1][i] = 0.0f; }
    for (int i = 0; i < n_sub; ++i) {
        const int st = i % D;
        const int rows_i = min(S, rows_here - i * S);
        mbar_wait(smem_u32(bars + st), (uint32_t)((i / D) & 1));
        const uint8_t* sb = ring + st * SG_STAGE_BYTES;
        for (int g = tr * 4; g < rows_i; g += 32) {             // groups of 4 consecutive rows
            float d[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t w = (g + u < rows_i) ? *reinterpret_cast<const uint32_t*>(sb + ((size_t)(g + u) * C + cp * 2) * 2) : 0u;
                d[u][0] = __uint_as_float(w << 16);
                d[u][1] = __uint_as_float(w & 0xffff0000u);
            }
            const int j0 = i * S + g;                             // xs column of the group's first row, tap 0
#pragma unroll
            for (int ci = 0; ci < CXM; ++ci) {
                if (ci < Cx) {
                    const float4 xa = *reinterpret_cast<const float4*>(xs + ci * pitch + j0);
                    const float2 xb = *reinterpret_cast<const float2*>(xs + ci * pitch + j0 + 4);
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
        __syncthreads();
        if (threadIdx.x == 0 && i + D < n_sub) issue(i + D);
    }
    float* red = reinterpret_cast<float*>(ring);                  // [8][C][nv]
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
#pragma unroll
        for (int t = 0; t < 8; ++t) sacc += red[(size_t)t * C * nv + i];
        pt[i] = sacc;
    }
}

// The same gradient as a tensor-core GEMM for Cx <= 8 and whole 64-row stages: dW^T is M = 64 couts x N = 3*Cx (padded to 24) taps
// with K = rows.  The CUDA-core kernel above is FMA-bound (2 x 21 FMAs per row and channel: 127 us at B = 256, L = 4096 against
// ~25 us of HBM time).  mma.sync.m16n8k16 (bf16 x bf16 -> fp32): the A operand d_raw^T comes straight out of the row-major
// stage with ldmatrix.trans (exact: d_raw is bf16), the B operand is the input window rounded to bf16 on the fly.
// Warp w owns k-step (w & 3) of every 64-row stage and couts [32 (w >> 2), +32).
__global__ void __launch_bounds__(256) wgrad_in_mma_kernel(const float* __restrict__ x, int Cx, int L, const bf16* __restrict__ d_raw,
                                                           float* __restrict__ partial, int rows_per_cta) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int C = 64, S = SG_STAGE_BYTES / (C * 2), D = SG_DEPTH;       // S = 64 rows per stage
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem;                                         // [D][8 KB]; reused for the final reduction
    const int pitch = rows_per_cta + 8;
    const int nv = Cx * 3;
    const int ring_bytes = max(D * SG_STAGE_BYTES, 4 * C * nv * 4);
    float* xs = reinterpret_cast<float*>(smem + ring_bytes);      // [Cx][pitch]: xs[ci][j] = x[ci][r0 - 4 + j] (aligned body at j = 4)
    uint64_t* bars = reinterpret_cast<uint64_t*>(xs + Cx * pitch);
    const int b = blockIdx.y, r0 = blockIdx.x * rows_per_cta;
    const int rows_here = min(rows_per_cta, L - r0);              // a multiple of 64 (launcher)
    const int n_sub = rows_here / S;
    const bf16* dbase = d_raw + ((size_t)b * L + r0) * C;
    auto issue = [&](int i) {
        const uint32_t bar = smem_u32(bars + (i % D));
        mbar_expect_tx(bar, SG_STAGE_BYTES);
        bulk_load(smem_u32(ring + (i % D) * SG_STAGE_BYTES), dbase + (size_t)i * S * C, SG_STAGE_BYTES, bar);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < D; ++s) mbar_init(smem_u32(bars + s), 1);
        fence_barrier_init();
        fence_proxy_async();
        for (int i = 0; i < D && i < n_sub; ++i) issue(i);
    }
    {   // input staging with vector loads, a thread's loads all in flight before its first store
        const int Q = rows_per_cta / 4;
        constexpr int UNR = 4;
        for (int i0 = threadIdx.x; i0 < Cx * Q; i0 += 256 * UNR) {
            float4 v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int i = i0 + u * 256;
                const int c = i / Q, l = r0 + 4 * (i % Q);
                v[u] = (i < Cx * Q && l < L) ? *reinterpret_cast<const float4*>(x + ((size_t)b * Cx + c) * L + l)
                                             : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int i = i0 + u * 256;
                if (i < Cx * Q) *reinterpret_cast<float4*>(xs + (i / Q) * pitch + 4 + 4 * (i % Q)) = v[u];
            }
        }
        if (threadIdx.x < 2 * Cx) {
            const int c = threadIdx.x >> 1, side = threadIdx.x & 1;
            const int l = side ? r0 + rows_per_cta : r0 - 1;
            xs[c * pitch + (side ? 4 + rows_per_cta : 3)] = (l >= 0 && l < L) ? x[((size_t)b * Cx + c) * L + l] : 0.0f;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const int kq = warp & 3, mh = warp >> 2;
    // B fragment column n = 8 nt + g  ->  (input channel, tap)  ->  offset into xs (-1: padding column)
    int boff[3];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
        const int kk = nt * 8 + g;
        boff[nt] = kk < nv ? (kk / 3) * pitch + kk % 3 + 3 : -1;
    }
    // ldmatrix.x4.trans: lane -> row (k) and 8-column block (m) of the four 8x8 matrices a0..a3 of one 16 x 16 A tile
    const int lk = (lane & 7) + 8 * (lane >> 4), lm = 8 * ((lane >> 3) & 1);
    float acc[2][3][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[m][nt][q] = 0.0f;
    for (int i = 0; i < n_sub; ++i) {
        const int st = i % D;
        mbar_wait(smem_u32(bars + st), (uint32_t)((i / D) & 1));
        const uint32_t sb = smem_u32(ring + st * SG_STAGE_BYTES);
        const int k0 = kq * 16;
        uint32_t bfr[3][2];
        const float* xw = xs + i * S + k0 + 2 * t4;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
            if (boff[nt] >= 0) {
                const float* xp = xw + boff[nt];
                bfr[nt][0] = pack_bf16x2(xp[0], xp[1]);
                bfr[nt][1] = pack_bf16x2(xp[8], xp[9]);
            } else {
                bfr[nt][0] = 0u; bfr[nt][1] = 0u;
            }
        }
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            uint32_t a0, a1, a2, a3;
            const uint32_t addr = sb + (uint32_t)(((k0 + lk) * C + 32 * mh + 16 * m + lm) * 2);
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr));
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[m][nt][0]), "+f"(acc[m][nt][1]), "+f"(acc[m][nt][2]), "+f"(acc[m][nt][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bfr[nt][0]), "r"(bfr[nt][1]));
        }
        __syncthreads();
        if (threadIdx.x == 0 && i + D < n_sub) issue(i + D);
    }
    float* red = reinterpret_cast<float*>(ring);                  // [4 k-steps][C][nv]
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int co = 32 * mh + 16 * m + g + 8 * (q >> 1), kk = 8 * nt + 2 * t4 + (q & 1);
                if (kk < nv) red[((size_t)kq * C + co) * nv + kk] = acc[m][nt][q];
            }
    __syncthreads();
    float* pt = partial + ((size_t)b * gridDim.x + blockIdx.x) * C * nv;
    for (int i = threadIdx.x; i < C * nv; i += 256)
        pt[i] = (red[i] + red[(size_t)C * nv + i]) + (red[(size_t)2 * C * nv + i] + red[(size_t)3 * C * nv + i]);
}

int g_wgrad_in_mma = 1;

// launcher used by gw_wgrad_in (backward.cu) for bf16, C = 64, L % 4 == 0; returns the number of partial rows written
int wgrad_in_stream(const float* x, int B, int Cx, int L, const void* d_raw, float* scratch, long scratch_elems, int* n_rows,
                    cudaStream_t st) {
    const int C = 64, rows = L < 512 ? ((L + 3) & ~3) : 512;
    const int n_rc = gw_cdiv(L, rows), nv = Cx * 3;
    GW_REQUIRE((long)B * n_rc * C * nv <= scratch_elems, "gw_wgrad_in: scratch too small");
    const int ring = SG_DEPTH * SG_STAGE_BYTES > 8 * C * nv * 4 ? SG_DEPTH * SG_STAGE_BYTES : 8 * C * nv * 4;
    const size_t smem = (size_t)ring + (size_t)Cx * (rows + 8) * 4 + 64;
    dim3 grid(n_rc, B);
    if (g_wgrad_in_mma && Cx <= 8 && L % 64 == 0 && rows % 64 == 0) {
        GW_CUDA(cudaFuncSetAttribute(wgrad_in_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GW_CUDA(cudaFuncSetAttribute(wgrad_in_mma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        GW_CUDA(gw_launch_pdl(wgrad_in_mma_kernel, grid, dim3(256), (size_t)(smem), st, x, Cx, L, (const bf16*)d_raw, scratch, rows));
        GW_LAUNCH_CHECK();
        *n_rows = B * n_rc;
        return GW_OK;
    }
#define WIS_GO(CXM)                                                                                                   \
    do {                                                                                                              \
        GW_CUDA(cudaFuncSetAttribute(wgrad_in_stream_kernel<CXM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        GW_CUDA(cudaFuncSetAttribute(wgrad_in_stream_kernel<CXM>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared)); \
        GW_CUDA(gw_launch_pdl(wgrad_in_stream_kernel<CXM>, grid, dim3(256), (size_t)(smem), st, x, Cx, L, (const bf16*)d_raw, scratch, rows));               \
    } while (0)
    if (Cx <= 4) WIS_GO(4); else if (Cx <= 8) WIS_GO(8); else WIS_GO(16);
#undef WIS_GO
    GW_LAUNCH_CHECK();
    *n_rows = B * n_rc;
    return GW_OK;
}

// ================================================================================================
// head conv (C = 64 -> 1, k = 3) + CFG combine + DDIM/DDPM update as a streaming kernel (see final_step_kernel in forward.cu
// for the math and the argument meaning).  A CTA owns FSS_TP output positions of one sample: the FSS_TP + 2 rows of h it
// needs go through the bulk-copy ring, every row leaves three tap dots in shared memory, the update runs at the end.
// ================================================================================================
#define FSS_TP 508                    // positions per CTA (multiple of 4: Philox quads stay aligned); + 2 halo rows = 8 stages
#include "step_update.cuh"
__global__ void __launch_bounds__(256) final_step_stream_kernel(const bf16* __restrict__ h, const float* __restrict__ net_a,
                                                                const float* __restrict__ net_b, int B, int Cx, int L,
                                                                const float* __restrict__ wf, const float* __restrict__ bf,
                                                                FssArgs p, const float* __restrict__ coef,
                                                                const int* __restrict__ step_ptr, const float* __restrict__ noise,
                                                                float* __restrict__ eps_out, float* __restrict__ x0_out) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int C = 64, S = 64, D = SG_DEPTH, ROWS = FSS_TP + 2;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem;                                           // [D][8 KB]
    float* pd = reinterpret_cast<float*>(smem + D * SG_STAGE_BYTES);  // [2 halves][3 taps][ROWS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(pd + 2 * 3 * ROWS);
    const int b = blockIdx.y, l0 = blockIdx.x * FSS_TP;
    const int step = step_ptr != nullptr ? *step_ptr : 0;
    const float* net_in = (step & 1) ? net_b : net_a;
    float* net_out = const_cast<float*>((step & 1) ? net_a : net_b);
    const int n_half = (p.mode == 1 && p.cfg_both) ? 2 : 1;
    // rows of h this CTA streams: [w0, w1) = [l0 - 1, l0 + FSS_TP + 1) clipped to the sample; local index r = row - (l0 - 1)
    const int w0 = max(l0 - 1, 0), w1 = min(l0 + FSS_TP + 1, L);
    const int n_rows = w1 - w0, n_sub = (n_rows + S - 1) / S;
    const int r_off = w0 - (l0 - 1);                                // 1 for the first CTA of a sample, else 0
    const int total = n_half * n_sub;
    auto issue = [&](int i) {
        const int hf = i / n_sub, j = i % n_sub;
        const int rows_i = min(S, n_rows - j * S);
        const uint32_t bar = smem_u32(bars + (i % D));
        mbar_expect_tx(bar, (uint32_t)rows_i * C * 2);
        bulk_load(smem_u32(ring + (i % D) * SG_STAGE_BYTES), h + ((size_t)(b + hf * B) * L + w0 + (size_t)j * S) * C,
                  (uint32_t)rows_i * C * 2, bar);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < D; ++s) mbar_init(smem_u32(bars + s), 1);
        fence_barrier_init();
        fence_proxy_async();
        for (int i = 0; i < D && i < total; ++i) issue(i);
    }
    for (int i = threadIdx.x; i < 2 * 3 * ROWS; i += 256) pd[i] = 0.0f;      // rows outside the sample contribute zero
    const int sx = threadIdx.x & 3, tr = threadIdx.x >> 2;          // 4 threads per row (16 channels each), 64 rows per pass
    float w0v[16], w1v[16], w2v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int c = sx * 16 + i;
        w0v[i] = wf[c * 3 + 0];
        w1v[i] = wf[c * 3 + 1];
        w2v[i] = wf[c * 3 + 2];
    }
    __syncthreads();
    for (int i = 0; i < total; ++i) {
        const int hf = i / n_sub, j = i % n_sub;
        const int rows_i = min(S, n_rows - j * S);
        mbar_wait(smem_u32(bars + (i % D)), (uint32_t)((i / D) & 1));
        const uint8_t* sb = ring + (i % D) * SG_STAGE_BYTES;
        {
            const int r = tr;
            float v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = 0.0f;
            if (r < rows_i) {
                float va[8], vb[8];
                ld8(reinterpret_cast<const bf16*>(sb + ((size_t)r * C + sx * 16) * 2), va);
                ld8(reinterpret_cast<const bf16*>(sb + ((size_t)r * C + sx * 16 + 8) * 2), vb);
#pragma unroll
                for (int q = 0; q < 8; ++q) { v[q] = va[q]; v[8 + q] = vb[q]; }
            }
            float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                d0 = fmaf(v[q], w0v[q], d0);
                d1 = fmaf(v[q], w1v[q], d1);
                d2 = fmaf(v[q], w2v[q], d2);
            }
#pragma unroll
            for (int o = 2; o > 0; o >>= 1) {
                d0 += __shfl_xor_sync(0xffffffffu, d0, o);
                d1 += __shfl_xor_sync(0xffffffffu, d1, o);
                d2 += __shfl_xor_sync(0xffffffffu, d2, o);
            }
            if (sx == 0 && r < rows_i) {
                const int rl = r_off + j * S + r;                   // local row index in [0, ROWS)
                float* q = pd + hf * 3 * ROWS;
                q[rl] = d0;
                q[ROWS + rl] = d1;
                q[2 * ROWS + rl] = d2;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && i + D < total) issue(i + D);
    }
    // ---- update: thread g finishes the 4 consecutive positions l0 + 4g .. l0 + 4g + 3 (one Philox call per quad)
    const float wx0 = wf[C * 3 + 0], wx1 = wf[C * 3 + 1], wx2 = wf[C * 3 + 2], bias = bf[0];
    const int g = threadIdx.x;
    if (g * 4 >= FSS_TP || l0 + g * 4 >= L) return;
    float cfv[10] = {0.0f, 1.0f, 1.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f};
    if (p.mode == 1) {
        const float* cf = coef + (size_t)step * 16;
#pragma unroll
        for (int i = 0; i < 10; ++i) cfv[i] = cf[i];
    }
    FssCoef cf;
    cf.c_s1mab = cfv[0]; cf.c_sab = cfv[1]; cf.c_sabp = cfv[2]; cf.c_dir = cfv[3]; cf.c_sig = cfv[4]; cf.c_w = cfv[5];
    cf.use = (int)cfv[6]; cf.last = (int)cfv[7]; cf.draw = (int)cfv[8]; cf.c_s1mab_cl = cfv[9];
    const bool need_z = p.mode == 1 && !cf.last && cf.c_sig > 0.0f;
    float z4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (need_z && noise == nullptr) {
        const unsigned long long sd = p.rng != nullptr ? p.rng[0] : p.seed;
        const long s0 = p.rng != nullptr ? (long)p.rng[1] : p.sample0;
        Philox::normal4(sd, (uint32_t)(s0 + b), (uint32_t)step + 1u, (uint32_t)((l0 + g * 4) >> 2), z4);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int pi = g * 4 + u, l = l0 + pi;
        if (l >= L) break;
        float outv0 = 0.0f, outv1 = 0.0f, xt_c = 0.0f;
        for (int hf = 0; hf < n_half; ++hf) {
            const float* xr = net_in + (size_t)(b + hf * B) * Cx * L;
            const float xm = l > 0 ? xr[l - 1] : 0.0f, xc = xr[l], xp = l + 1 < L ? xr[l + 1] : 0.0f;
            const float* q = pd + hf * 3 * ROWS;
            const int r = pi + 1;
            float acc = q[r - 1] + q[ROWS + r] + q[2 * ROWS + r + 1];
            acc += fmaf(xm, wx0, fmaf(xc, wx1, xp * wx2));
            if (hf == 0) { outv0 = acc + bias; xt_c = xc; } else { outv1 = acc + bias; }
        }
        if (p.mode == 0) {
            eps_out[(size_t)b * L + l] = outv0;
            continue;
        }
        fss_update(p, cf, outv0, outv1, xt_c, z4[u], noise, net_out, n_half, b, B, Cx, L, l, eps_out, x0_out);
    }
}

// The same head + update from the three dot products per position that the last decoder's fused kernel leaves behind
// (gw_conv_gn2, dots [Bn, L, 4] fp32 = (sum_c h[l,c] w[c,0], .. w[c,1], .. w[c,2], 0)): eps_hat[l] = d0[l-1] + d1[l] + d2[l+1] + the x_t
// taps + bias (models.py:230).  16 B per position instead of the 128 B row of h; a thread owns 4 positions (one Philox quad).
__global__ void __launch_bounds__(256) final_step_dots_kernel(const float4* __restrict__ dots, const float* __restrict__ net_a,
                                                              const float* __restrict__ net_b, int B, int Cx, int L,
                                                              const float* __restrict__ wf, const float* __restrict__ bf, int C,
                                                              FssArgs p, const float* __restrict__ coef,
                                                              int* __restrict__ step_ptr, const float* __restrict__ noise,
                                                              float* __restrict__ eps_out, float* __restrict__ x0_out) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.y, l4 = (blockIdx.x * 256 + threadIdx.x) * 4;
    pdl_wait();                                                       // the dots and the step counter come from the previous kernels
    pdl_launch_dependents();
    const int step = step_ptr != nullptr ? *step_ptr : 0;
    if (l4 < L) {
    const float* net_in = (step & 1) ? net_b : net_a;
    float* net_out = const_cast<float*>((step & 1) ? net_a : net_b);
    const int n_half = (p.mode == 1 && p.cfg_both) ? 2 : 1;
    const float wx0 = wf[C * 3 + 0], wx1 = wf[C * 3 + 1], wx2 = wf[C * 3 + 2], bias = bf[0];
    float cfv[10] = {0.0f, 1.0f, 1.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f};
    if (p.mode == 1) {
        const float* cfp = coef + (size_t)step * 16;
#pragma unroll
        for (int i = 0; i < 10; ++i) cfv[i] = cfp[i];
    }
    FssCoef cf;
    cf.c_s1mab = cfv[0]; cf.c_sab = cfv[1]; cf.c_sabp = cfv[2]; cf.c_dir = cfv[3]; cf.c_sig = cfv[4]; cf.c_w = cfv[5];
    cf.use = (int)cfv[6]; cf.last = (int)cfv[7]; cf.draw = (int)cfv[8]; cf.c_s1mab_cl = cfv[9];
    const bool need_z = p.mode == 1 && !cf.last && cf.c_sig > 0.0f;
    float z4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (need_z && noise == nullptr) {
        const unsigned long long sd = p.rng != nullptr ? p.rng[0] : p.seed;
        const long s0 = p.rng != nullptr ? (long)p.rng[1] : p.sample0;
        Philox::normal4(sd, (uint32_t)(s0 + b), (uint32_t)step + 1u, (uint32_t)(l4 >> 2), z4);
    }
    const int n = min(4, L - l4);
    float ov[2][4], xc4[4];
    for (int hf = 0; hf < n_half; ++hf) {
        const float4* dr = dots + (size_t)(b + hf * B) * L;
        const float* xr = net_in + (size_t)(b + hf * B) * Cx * L;
        float dm = l4 > 0 ? dr[l4 - 1].x : 0.0f;                   // tap 0 of the previous position
        float xm = l4 > 0 ? xr[l4 - 1] : 0.0f;
        float4 dc = dr[l4];
        float xc = xr[l4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (u < n) {
                const int l = l4 + u;
                const bool nx = l + 1 < L;
                const float4 dn = nx ? dr[l + 1] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                const float xp = nx ? xr[l + 1] : 0.0f;
                float acc = dm + dc.y + dn.z;
                acc += fmaf(xm, wx0, fmaf(xc, wx1, xp * wx2));
                ov[hf][u] = acc + bias;
                if (hf == 0) xc4[u] = xc;
                dm = dc.x; dc = dn; xm = xc; xc = xp;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (u >= n) break;
        const int l = l4 + u;
        if (p.mode == 0) {
            eps_out[(size_t)b * L + l] = ov[0][u];
            continue;
        }
        fss_update(p, cf, ov[0][u], n_half == 2 ? ov[1][u] : 0.0f, xc4[u], z4[u], noise, net_out, n_half, b, B, Cx, L, l, eps_out, x0_out);
    }
    }
    if (p.advance != nullptr) {
        // every CTA has read *step_ptr before it arrives here: the last one out advances the step counter for the next launch
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int done = atomicAdd(p.advance, 1u);
            if (done == gridDim.x * gridDim.y - 1) {
                *p.advance = 0u;
                *step_ptr = step + 1;
            }
        }
    }
}

int final_step_dots(const void* dots, const float* net_a, const float* net_b, int B, int Cx, int L, int C, const float* wf,
                    const float* bf, const gw_step_params* p, const float* coef, const int* step_ptr, const float* noise,
                    float* eps_out, float* x0_out, cudaStream_t st) {
    FssArgs a;
    a.mode = p->mode; a.cfg_both = p->cfg_both; a.selfcond = p->selfcond; a.pred_x0 = p->pred_x0;
    a.eps_scale = p->eps_scale; a.dc_weight = p->dc_weight; a.y_dc = p->y_dc; a.seed = p->seed; a.sample0 = p->sample0; a.rng = p->rng;
    a.advance = (p->mode == 1 && step_ptr != nullptr) ? (unsigned int*)p->advance : nullptr;
    dim3 grid(gw_cdiv(L, 1024), B);
    GW_CUDA(gw_launch_pdl(final_step_dots_kernel, grid, dim3(256), (size_t)0, st, (const float4*)dots, net_a, net_b ? net_b : net_a, B,
                          Cx, L, wf, bf, C, a, coef, const_cast<int*>(step_ptr), noise, eps_out, x0_out));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// called by gw_final_step (forward.cu) for bf16, C = 64
int final_step_stream(const void* h, const float* net_a, const float* net_b, int B, int Cx, int L, const float* wf, const float* bf,
                      const gw_step_params* p, const float* coef, const int* step_ptr, const float* noise, float* eps_out,
                      float* x0_out, cudaStream_t st) {
    FssArgs a;
    a.mode = p->mode; a.cfg_both = p->cfg_both; a.selfcond = p->selfcond; a.pred_x0 = p->pred_x0;
    a.eps_scale = p->eps_scale; a.dc_weight = p->dc_weight; a.y_dc = p->y_dc; a.seed = p->seed; a.sample0 = p->sample0; a.rng = p->rng;
    a.advance = nullptr;
    const size_t smem = (size_t)SG_DEPTH * SG_STAGE_BYTES + (size_t)2 * 3 * (FSS_TP + 2) * 4 + 64;
    GW_CUDA(cudaFuncSetAttribute(final_step_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GW_CUDA(cudaFuncSetAttribute(final_step_stream_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    dim3 grid(gw_cdiv(L, FSS_TP), B);
    GW_CUDA(gw_launch_pdl(final_step_stream_kernel, grid, dim3(256), (size_t)(smem), st, (const bf16*)h, net_a, net_b ? net_b : net_a, B, Cx, L, wf, bf, a, coef, step_ptr,
                                                      noise, eps_out, x0_out));
    GW_LAUNCH_CHECK();
    return GW_OK;
}
