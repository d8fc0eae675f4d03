// Fused conv block: Conv1d(k=3) -> GroupNorm(8) -> SiLU -> + cond 1x1 conv -> FiLM (-> avg_pool) in ONE kernel (sm_100a).
// Included at the end of conv_tc.cu (shares TcParams / build_params / the tensor-map helpers).
//
// GroupNorm needs the statistics of a whole sample before any output element can be produced, and the fp32 conv output of
// one sample-layer is 0.5-1 MB -- more than one SM's tensor memory (256 KB).  So a GROUP of G persistent CTAs (G = sample
// accumulators / (128 lanes x 256 columns): 8 at L = 4096, 4 for the bottleneck) owns a sample: every CTA keeps its
// 128 x 256 fp32 slice of the conv output in TMEM, reduces its GroupNorm partial sums (pass 1 over TMEM), exchanges 16
// floats with the other CTAs of the group through 8-byte {value, epoch} packets in global memory (no fence, no reset: the
// epoch is a per-launch counter, so a stale packet never matches), and then normalises / activates / modulates straight
// out of TMEM (pass 2) into swizzled staging tiles that leave through TMA stores.  The raw conv output is never written
// and re-read (inference), or written once for the backward pass (training) -- the unfused path writes it and reads it
// back (gw_conv_tc + gw_gn_apply).  TMEM holds two samples per CTA, so the MMA warp runs one sample ahead of the epilogue
// and the exchange latency (~1-2 us) is hidden behind the next sample's MMAs.
//
// All CTAs of a group spin on each other's packets, so the whole grid must be co-resident: grid = G x n_groups <= #SMs and
// one CTA per SM (the shared-memory footprint guarantees it).
#pragma once
#include "xchg.cuh"

#define CGN_MAX_G 32
#define CGN_NCA_MAX 8
#define CGN_EPI_GROUPS 2       // epilogue warpgroups (8 warps each) working on alternate samples
#define CGN_THREADS (64 + CGN_EPI_GROUPS * 256 + 32)     // TMA warp, MMA warp, epilogue warpgroups, chain-signalling warp
// 1: the 16 epilogue warps form ONE group (every warp owns one 32x64 chunk of every sample); 0: two groups of 8 warps on
// alternate samples (two chunks per warp)
#ifndef CGN_ONE_GROUP
#define CGN_ONE_GROUP 0
#endif

struct GnFuseArgs {
    const float* gn_w;
    const float* gn_b;
    const float* cond;        // [B, Lpos, Cc] fp32 or NULL
    const float* wc;          // [C, Cc]
    const float* bc;          // [C]
    const float* film;
    const int* step_ptr;
    float* stats_out;         // [B, 8, 2] (mean, rstd) or NULL
    unsigned long long* xchg; // [B][CGN_MAX_G][16] packets
    unsigned int* ctrl;       // [0] epoch of the previous launch, [1] CTAs finished
    long film_b_stride, film_step_stride;
    int film_off, Cc, G, n_groups, B, Lpos, write_raw, pair;
    const float* head_w;      // HEAD kernels: final.weight [C+1, 3] (models.py:230), else NULL
    float* head_dots;         // HEAD kernels: [B, Lpos, 4] fp32: (sum_c out[l,c] w[c,0], .. w[c,1], .. w[c,2], 0)
    int store_out;            // 0: the activated tensor itself is not needed (only the head dots leave the kernel)
    int pair2;                // 1: CTA pairs (cluster of 2, tcgen05 cta_group::2): ONE M = 256 MMA per pair, each CTA stages its own
                              //    A rows and HALF of every weight tile -- halves the weight bytes an SM pulls through the L2
    int wbox;                 // rows of the weight TMA box (64, or 32 when a pair splits a 64-row weight tile)
    // ---- layer chaining (sampler): the NEXT layer's kernel is launched with programmatic stream serialization and starts on
    // the SMs this kernel's early-finishing groups free; it orders itself per SAMPLE through these flags instead of waiting for
    // the whole grid: no fill / drain bubble between layers and the last, partly filled round overlaps the next layer's work
    unsigned int* done_flag;  // [B][CGN_MAX_G]: CTA j of a sample's group writes V once all ITS stores of sample b have completed; NULL: off
    const unsigned int* prev_flag;   // the producing layer's done_flag, or NULL: plain griddepcontrol.wait ordering
    int prev_G;               // CTAs per sample of the producing layer
    const int* serial_ptr;    // device chain serial; V = serial * 4096 + step + 1 is unique per (chain, reverse step)
    int dbg_mode;             // tools only (-DCGN_ABLATE builds): 1 no tanh, 2 no pack, 4 no stmatrix, 8 no TMA stores, 16 no pooling, 32 no TMEM load
    long long* dbg;           // tools only: [CTA][16 samples][8] clock64 stamps of CTA phases (NULL in production)
    unsigned long long* tdbg; // tools only: [CTA][2] %globaltimer at CTA start / end of this launch (NULL in production)
};
#define CGN_STAMP(k)                                                                                     \
    do {                                                                                                 \
        if (F.dbg != nullptr && it < 16) F.dbg[((size_t)blockIdx.x * 16 + it) * 8 + (k)] = clock64();    \
    } while (0)

static __device__ __noinline__ void xchg_timeout(int b, int src) {
    printf("gwb200 conv_gn kernel: statistics exchange timed out (block %d sample %d source %d)\n", blockIdx.x, b, src);
    __trap();
}

// The epilogue is written as SMALL LOOPS on purpose: two warps per scheduler cannot hide instruction-fetch stalls, and a fully
// unrolled per-sample body (~3500 instructions) ran every phase at ~1/3 of the issue rate it reaches once it fits the
// instruction cache (measured with the clock64 stamps of tools/cgn_timeline.py).
//
// Pass 2 of 16 columns of a [32 rows x 64 columns] warp chunk: TMEM fragments -> GN affine -> SiLU -> FiLM/cond -> bf16 ->
// swizzled staging (stmatrix), and (POOL) the 2:1 row average into a second staging tile.  Thread t owns, per 8-column block
// k, the column pair 8k + 2(t%4) of rows t/4 + 8m (m = 0..3), so per-column coefficients are fetched once per 8 elements.
//   ab[pair] = {A.x, A.y, B.x, B.y}: h = A*acc + B is HALF the GroupNorm output (conv bias folded in), silu = h + h*tanh(h)
//   ge[pair] = {G.x, G.y, E.x, E.y}: out = silu*G + E + sum_j W_j * cond_j,  G = 1+gamma_t, E = bc*G + beta_t, W_j = wc_j*G
//   cd[m][j]: cond channel j at this thread's row of row-group m;  kb: which 16 columns of the chunk
//   HEAD: hw[pair][tap] = head-conv weights of the column pair; hd[m][tap] += out * w (summed over the chunk's 64 channels by the
//   caller); st_out = 0 skips the staging of the activated tile
template <int NCA, bool HAS_COND, bool POOL, bool HEAD = false>
__device__ __forceinline__ void gn_cols16(uint32_t taddr, int kb, const ulonglong2* __restrict__ ab, const ulonglong2* __restrict__ ge,
                                          const unsigned long long* __restrict__ w, const float (&cd)[4][NCA], uint32_t stg,
                                          uint32_t stgp, int lane, int dm = 0, const unsigned long long* __restrict__ hw = nullptr,
                                          unsigned long long (*hd)[3] = nullptr, bool st_out = true) {
    const unsigned long long half2 = pkf2(0.5f, 0.5f);
    const int tq = lane & 3, tr = lane >> 2;
    // coefficient loads first: the TMEM load below is an asm volatile with a memory clobber, nothing moves across it, and the
    // LDS -> FFMA2 latency would otherwise sit exposed at the head of every block
    ulonglong2 c_ab2[2], c_ge2[2];
    unsigned long long wv2[2][NCA];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        const int pi = 4 * (2 * kb + kk) + tq;
        c_ab2[kk] = ab[pi];
        c_ge2[kk] = ge[pi];
        if (HAS_COND) {
#pragma unroll
            for (int jj = 0; jj < NCA; ++jj) wv2[kk][jj] = w[pi * NCA + jj];
        }
    }
    uint32_t v[16];
#ifdef CGN_ABLATE
    if (dm & 32) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (uint32_t)(lane + i + kb) << 20;
    } else
#endif
    tmem_ld_frag16(taddr + (uint32_t)(kb * 16), v);
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        const int k = 2 * kb + kk;
        const ulonglong2 c_ab = c_ab2[kk];
        const ulonglong2 c_ge = c_ge2[kk];
        unsigned long long wv[NCA];
#pragma unroll
        for (int jj = 0; jj < NCA; ++jj) wv[jj] = HAS_COND ? wv2[kk][jj] : 0ull;
        unsigned long long o[4];
        uint32_t r[4];
        unsigned long long hwk[3] = {0ull, 0ull, 0ull};
        if (HEAD) {
            const int pi = 4 * k + tq;
            hwk[0] = hw[pi * 3 + 0]; hwk[1] = hw[pi * 3 + 1]; hwk[2] = hw[pi * 3 + 2];
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int i0 = 8 * (m >> 1) + 4 * kk + 2 * (m & 1);
            const unsigned long long h = ffma2(pk2(v[i0], v[i0 + 1]), c_ab.x, c_ab.y);
            float h0, h1, t0, t1;
            upk2(h, h0, h1);
#ifdef CGN_ABLATE
            if (dm & 1) { t0 = h0; t1 = h1; } else
#endif
            {
                asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
            }
            o[m] = ffma2(ffma2(h, pkf2(t0, t1), h), c_ge.x, c_ge.y);
            if (HAS_COND) {
#pragma unroll
                for (int jj = 0; jj < NCA; ++jj) o[m] = ffma2(wv[jj], pkf2(cd[m][jj], cd[m][jj]), o[m]);
            }
            if (HEAD) {
#pragma unroll
                for (int tp = 0; tp < 3; ++tp) hd[m][tp] = ffma2(o[m], hwk[tp], hd[m][tp]);
            }
            float lo, hi;
            upk2(o[m], lo, hi);
#ifdef CGN_ABLATE
            if (dm & 2) r[m] = __float_as_uint(lo) ^ __float_as_uint(hi); else
#endif
            r[m] = pack_bf16x2(lo, hi);
        }
#ifdef CGN_ABLATE
        if (!(dm & 4))
#endif
        if (!HEAD || st_out)
        stmatrix_x4(stg + (uint32_t)lane * 128 + (uint32_t)((k ^ (lane & 7)) * 16), r[0], r[1], r[2], r[3]);
#ifdef CGN_ABLATE
        if (dm & 16) continue;
#endif
        if (POOL) {
            // rows 8m + tr and 8m + (tr ^ 1) live in lanes t and t ^ 4.  The even lane pools row-groups 0, 1 and the odd one 3, 2
            // (each sends the partner the two values it needs): pooled rows tr/2 + {0, 4} resp. {12, 8} -- different swizzle
            // phases, so the two lanes' stores do not collide.
            const bool odd = (tr & 1) != 0;
            const unsigned long long ra = __shfl_xor_sync(0xffffffffu, odd ? o[0] : o[3], 4);
            const unsigned long long rb = __shfl_xor_sync(0xffffffffu, odd ? o[1] : o[2], 4);
            float lo, hi;
            upk2(fmul2(fadd2(odd ? o[3] : o[0], ra), half2), lo, hi);
            const uint32_t pa = pack_bf16x2(lo, hi);
            upk2(fmul2(fadd2(odd ? o[2] : o[1], rb), half2), lo, hi);
            const uint32_t pb = pack_bf16x2(lo, hi);
            const int p1 = (odd ? 12 : 0) + (tr >> 1), p2 = (odd ? 8 : 4) + (tr >> 1);
            const uint32_t d1 = stgp + (uint32_t)p1 * 128 + (uint32_t)((k ^ (p1 & 7)) * 16) + (uint32_t)tq * 4;
            const uint32_t d2 = stgp + (uint32_t)p2 * 128 + (uint32_t)((k ^ (p2 & 7)) * 16) + (uint32_t)tq * 4;
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(d1), "r"(pa) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(d2), "r"(pb) : "memory");
        }
    }
}

// Pass 1 of one warp chunk: per-LANE GroupNorm sums of (acc + bias) (the caller reduces across the warp), and (store != 0)
// the bf16 conv output into a swizzled staging tile for the backward pass.
template <int CG_LOG2>
__device__ __forceinline__ void stat_chunk64(uint32_t taddr, const float* sbias, bool row_ok, float (&sv)[2 * (64 >> CG_LOG2)],
                                             bool store, uint32_t stg, int lane) {
    constexpr int NG = 64 >> CG_LOG2;
    constexpr int BG = (1 << CG_LOG2) / 8;                    // 8-column blocks per group
    const int tq = lane & 3;
    unsigned long long a1[NG], a2[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) { a1[g] = 0ull; a2[g] = 0ull; }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld_frag32(taddr + (uint32_t)(half * 32), v);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int k = half * 4 + kk;
            const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(sbias + 8 * k + 2 * tq);
            unsigned long long x[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int i0 = 16 * (m >> 1) + 4 * kk + 2 * (m & 1);
                x[m] = fadd2(pk2(v[i0], v[i0 + 1]), bb);
                a1[k / BG] = fadd2(a1[k / BG], x[m]);
                a2[k / BG] = ffma2(x[m], x[m], a2[k / BG]);
            }
            if (store) {
                uint32_t r[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    float lo, hi;
                    upk2(x[m], lo, hi);
                    r[m] = pack_bf16x2(lo, hi);
                }
                stmatrix_x4(stg + (uint32_t)lane * 128 + (uint32_t)((k ^ (lane & 7)) * 16), r[0], r[1], r[2], r[3]);
            }
        }
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        float lo, hi;
        upk2(a1[g], lo, hi);
        sv[2 * g] = row_ok ? lo + hi : 0.0f;
        upk2(a2[g], lo, hi);
        sv[2 * g + 1] = row_ok ? lo + hi : 0.0f;
    }
}
// warp-reduce the per-lane sums of stat_chunk64 into this warp's [8 groups][2] slots
template <int CG_LOG2>
__device__ __forceinline__ void stat_reduce(float (&sv)[2 * (64 >> CG_LOG2)], float* stat, int grp0, int lane) {
    constexpr int NV = 2 * (64 >> CG_LOG2);
    const float tot = warp_reduce_multi<NV>(sv, lane);
    constexpr int LOW = NV == 16 ? 2 : (NV == 8 ? 4 : 8);   // lanes sharing one value
    if ((lane & (LOW - 1)) == 0) stat[grp0 * 2 + lane / LOW] += tot;
}

// CG_LOG2 = log2(channels per group) (3, 4, 5 <=> C = 64, 128, 256); MT = row tiles per CTA (MT * bn = 256 TMEM columns);
// CC = cond channels (1, 5, or -1: any Cc in [0, 8] with zero-padded weights); POOL: also write avg_pool1d(out, 2).
// HEAD (Cout = 64 decoder in pair space only): the head conv final(cat[h, x_t]) reads nothing but three dot products per position
// of this block's output, so the epilogue forms them from the fp32 values it already holds (gw_final_step then runs on 16 B per
// position instead of streaming the 128 B row back in).
// PAIR2: CTA pairs (cluster of 2, tcgen05 cta_group::2), a compile-time flavour: a kernel that contains .2CTA instructions can
// only be launched with an even cluster size.
// ONE_T: the 16 epilogue warps form ONE group that works on every sample (one 32 x 64 chunk per warp) instead of two warpgroups on
// alternate samples -- the flavour for launches in which a CTA sees a single sample (batch <= number of CTA groups: the second
// warpgroup would otherwise idle); +6 % on the B = 8 chain, -2 % at B = 256 (profiles/r02_one_group.md).
template <int CG_LOG2, int MT, int CC, bool POOL, bool HEAD = false, bool PAIR2 = false, bool ONE_T = (CGN_ONE_GROUP != 0)>
__global__ void __launch_bounds__(CGN_THREADS, 1)
conv_gn_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
               const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_out,
               const __grid_constant__ CUtensorMap tm_raw, const __grid_constant__ CUtensorMap tm_pool,
               const __grid_constant__ TcParams P, const __grid_constant__ GnFuseArgs F, const float* __restrict__ bias) {
    constexpr int NCA = CC > 0 ? CC : CGN_NCA_MAX;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int SA = P.sa, SB = P.sb, NBUF = P.nbuf;
    // B ring slot: the whole weight tile of a segment, or (pairs) the half this CTA stages -- half the bytes per slot buys a
    // ring twice as deep, i.e. twice as many segments in flight against the L2 -> shared-memory latency
    const uint32_t b_bytes = (uint32_t)P.bn * TC_BLOCK_K * (PAIR2 ? 1 : 2);
    const uint32_t sA = base;                                           // [SA][MT][A_SLOT]
    const uint32_t sB = sA + (uint32_t)SA * MT * TC2_A_SLOT;            // [SB][b_bytes]
    constexpr uint32_t EW = 8u * CGN_EPI_GROUPS;                        // epilogue warps
    const uint32_t sStage = sB + (uint32_t)SB * b_bytes;                // [EW warps][NBUF][4096]
    const uint32_t sPool = sStage + EW * NBUF * 4096u;                  // [EW warps][NBUF][2048] (POOL)
    const uint32_t sMisc = sPool + (POOL ? EW * NBUF * 2048u : 0u);
    uint8_t* misc = smem_raw + (sMisc - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(misc);                 // a_full[SA] a_empty[SA] b_full[SB] b_empty[SB] acc_full[2] acc_empty[2] sig[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 8 * 28);
    float* s_bias = reinterpret_cast<float*>(misc + 256);               // [256]
    // per epilogue warpgroup: s_stat [8 warps][8 groups][2] | s_x [G][16] | s_abf, s_gef [128 column pairs][4] | s_wf [128][NCA][2]
    constexpr int GRP_FLOATS = 256 + CGN_MAX_G * 16 + 512 + 512 + 256 * NCA;
    float* s_grp0 = s_bias + 256;
    unsigned long long* s_hw = reinterpret_cast<unsigned long long*>(s_grp0 + CGN_EPI_GROUPS * GRP_FLOATS);   // HEAD: [32 pairs][3 taps]
    auto a_full = [&](int i) { return smem_u32(bars + i); };
    auto a_empty = [&](int i) { return smem_u32(bars + SA + i); };
    auto b_full = [&](int i) { return smem_u32(bars + 2 * SA + i); };
    auto b_empty = [&](int i) { return smem_u32(bars + 2 * SA + SB + i); };
    auto acc_full = [&](int i) { return smem_u32(bars + 2 * SA + 2 * SB + i); };
    auto acc_empty = [&](int i) { return smem_u32(bars + 2 * SA + 2 * SB + 2 + i); };
    // chaining: monotonic per-warpgroup arrival counters "my stores of a sample completed" (an mbarrier's phase parity would alias:
    // nothing stops the epilogue from completing two samples before the signalling thread has looked)
    auto sig_cnt = [&](int i) { return smem_u32(bars + 2 * SA + 2 * SB + 4 + i); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = F.G;
    const uint32_t cta_rank = PAIR2 ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0u;
    const int grp = blockIdx.x / G, j_cta = blockIdx.x % G;
    // pairs: CTAs (2c, 2c+1) of a group share the weights (same n_tile) and own adjacent row slices
    const int n_tile = PAIR2 ? (j_cta >> 1) % P.n_tiles : j_cta % P.n_tiles;
    const int ms = PAIR2 ? ((j_cta >> 1) / P.n_tiles) * 2 + (j_cta & 1) : j_cta / P.n_tiles;
    const int row0 = ms * (MT * TC_BLOCK_M);
    const int n_seg = P.n_seg[n_tile];

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_a0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_a1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_out) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_raw) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_pool) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < SA; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
        for (int i = 0; i < SB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
        // pairs: the leader's MMA thread waits until BOTH epilogues have drained an accumulator stage
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), PAIR2 ? 2 : 1); }
        bars[2 * SA + 2 * SB + 4] = 0ull;
        bars[2 * SA + 2 * SB + 5] = 0ull;
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR2) tmem_alloc_2sm(smem_u32(tmem_slot), 512);
        else tmem_alloc(smem_u32(tmem_slot), 512);
    }
    for (int i = threadIdx.x; i < P.bn; i += blockDim.x) s_bias[i] = bias ? bias[(n_tile * P.bn + i) & (P.cout - 1)] : 0.0f;
    if (HEAD) {
        for (int i = threadIdx.x; i < 32 * 3; i += blockDim.x) {
            const int pi = i / 3, tp = i % 3;
            s_hw[i] = pkf2(F.head_w[(2 * pi) * 3 + tp], F.head_w[(2 * pi + 1) * 3 + tp]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR2) cluster_sync_all();          // the peer's barriers are initialised before any remote arrive / TMA completion
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above touched parameters only; the activations, the exchange buffer and the step counter belong to the previous
    // kernel of the stream (PDL, common.cuh).  Chained launches order themselves per sample (producer warp below): everything
    // else this kernel reads was written before the previous non-chained kernel of the stream completed.
    if (F.prev_flag == nullptr) pdl_wait();
    pdl_launch_dependents();
    if (F.tdbg != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        F.tdbg[blockIdx.x * 2] = t;
    }
    const unsigned int chain_v = F.serial_ptr != nullptr
                                     ? (unsigned int)(*F.serial_ptr) * 4096u + (unsigned int)(F.step_ptr != nullptr ? *F.step_ptr : 0) + 1u
                                     : 0u;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t ia = 0, pa = 0, ib = 0, pb = 0;
            const int wbox = F.wbox;
            unsigned int pfl[8];                          // chain flags of the sample about to be loaded (prev_G <= 8 used)
            auto load_flags = [&](int bb) {
                const unsigned int* fp = F.prev_flag + (size_t)bb * CGN_MAX_G;
                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(pfl[0]), "=r"(pfl[1]), "=r"(pfl[2]), "=r"(pfl[3]) : "l"(fp) : "memory");
                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(pfl[4]), "=r"(pfl[5]), "=r"(pfl[6]), "=r"(pfl[7]) : "l"(fp + 4) : "memory");
            };
            if (F.prev_flag != nullptr && grp < F.B) load_flags(grp);
            for (int b = grp; b < F.B; b += F.n_groups) {
                if (F.prev_flag != nullptr) {
                    // chained launch: sample b of the producing layer (and, transitively, of every layer before it) is complete
                    // (every CTA of its group has published V for b).  Relaxed accesses, as for the statistics packets (xchg.cuh): a
                    // flag is stored only after the bulk stores it covers have COMPLETED (performed at the L2), the loads below are
                    // issued only after the flag was read from the L2, and the tensor data never passes through an L1.
                    // The flags of a sample are read with two 16-byte loads; those of the NEXT sample are requested right away, so
                    // in steady state the check costs no L2 round trip.
                    const long long t0 = clock64();
                    for (;;) {
                        bool ok = true;
#pragma unroll
                        for (int j = 0; j < 8; ++j) ok = ok && (j >= F.prev_G || pfl[j] == chain_v);
                        if (ok) break;
                        __nanosleep(64);
                        load_flags(b);
                        if (clock64() - t0 > 4000000000LL) xchg_timeout(b, -1);
                    }
                    if (b + F.n_groups < F.B) load_flags(b + F.n_groups);
                    asm volatile("fence.proxy.async;" ::: "memory");       // the TMA loads below must not pass the polls
                }
                for (int s = 0; s < n_seg; ++s) {
                    const TcSeg sg = P.seg[n_tile][s];
                    if (sg.a_new) {
                        mbar_wait(a_empty(ia), pa ^ 1);
                        if (!PAIR2) {
                            mbar_expect_tx(a_full(ia), MT * TC2_A_BYTES);
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt)
                                tma_load_3d(sA + (ia * MT + mt) * TC2_A_SLOT, sg.src ? &tm_a1 : &tm_a0, a_full(ia), sg.col,
                                            row0 + mt * TC_BLOCK_M + sg.load_shift, b);
                        } else {
                            // both CTAs' boxes are counted on the LEADER's barrier (its own smem holds this CTA's rows)
                            if (leader) mbar_expect_tx(a_full(ia), 2 * MT * TC2_A_BYTES);
                            const uint32_t lbar = mapa_u32(a_full(ia), 0);
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt)
                                tma_load_3d_2sm(sA + (ia * MT + mt) * TC2_A_SLOT, sg.src ? &tm_a1 : &tm_a0, lbar, sg.col,
                                                row0 + mt * TC_BLOCK_M + sg.load_shift, b);
                        }
                        if (++ia == (uint32_t)SA) { ia = 0; pa ^= 1; }
                    }
                    mbar_wait(b_empty(ib), pb ^ 1);
                    const int wrow = n_tile * P.bn + sg.n_off;
                    if (!PAIR2) {
                        mbar_expect_tx(b_full(ib), (uint32_t)sg.n_cnt * TC_BLOCK_K * 2);
                        for (int j = 0; j < sg.n_cnt; j += 64)
                            tma_load_2d(sB + ib * b_bytes + (uint32_t)j * 128, &tm_w, b_full(ib), sg.wk, wrow + j);
                    } else {
                        // this CTA stages weight rows [rank * n/2, (rank + 1) * n/2) at the start of its slot: the MMA reads B rows
                        // [0, n/2) from the leader's shared memory and [n/2, n) from the peer's, at the same offset
                        if (leader) mbar_expect_tx(b_full(ib), (uint32_t)sg.n_cnt * TC_BLOCK_K * 2);
                        const uint32_t lbar = mapa_u32(b_full(ib), 0);
                        const int half = sg.n_cnt >> 1;
                        const int r0 = wrow + (int)cta_rank * half;
                        for (int j = 0; j < half; j += wbox)
                            tma_load_2d_2sm(sB + ib * b_bytes + (uint32_t)j * 128, &tm_w, lbar, sg.wk, r0 + j);
                    }
                    if (++ib == (uint32_t)SB) { ib = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (pairs: the leader CTA only) =====================
        if (lane == 0 && leader) {
            uint32_t ia = 0, pa = 0, ib = 0, pb = 0, cur_a = 0;
            const uint64_t desc_hi = make_sw128_desc(0, 0);
            int it = 0;
            for (int b = grp; b < F.B; b += F.n_groups, ++it) {
                const int as = it & 1;
                mbar_wait(acc_empty(as), (((uint32_t)(it >> 1)) & 1u) ^ 1u);
                tc_fence_after();
                CGN_STAMP(6);
                const uint32_t acc0 = tmem_base + (uint32_t)(as * 256);
                for (int s = 0; s < n_seg; ++s) {
                    const TcSeg sg = P.seg[n_tile][s];
                    if (sg.a_new) {
                        cur_a = ia;
                        mbar_wait(a_full(ia), pa);
                        if (++ia == (uint32_t)SA) { ia = 0; pa ^= 1; }
                    }
                    mbar_wait(b_full(ib), pb);
                    tc_fence_after();
                    const uint32_t idesc = make_idesc(PAIR2 ? 2 * TC_BLOCK_M : TC_BLOCK_M, (uint32_t)sg.n_cnt);
                    const uint64_t bd0 = desc_hi | (uint64_t)(((sB + ib * b_bytes) & 0x3FFFFu) >> 4);
                    const uint32_t acc_first = s > 0 ? 1u : 0u;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint32_t a0 = sA + (cur_a * MT + mt) * TC2_A_SLOT + (uint32_t)sg.desc_row * 128u;
                        const uint64_t ad0 = desc_hi | (uint64_t)((a0 & 0x3FFFFu) >> 4);
                        const uint32_t d = acc0 + (uint32_t)(mt * P.bn + sg.n_off);
                        if (PAIR2) {
                            umma_bf16_2sm(d, ad0, bd0, idesc, acc_first);
                            umma_bf16_2sm(d, ad0 + 2, bd0 + 2, idesc, 1u);
                            umma_bf16_2sm(d, ad0 + 4, bd0 + 4, idesc, 1u);
                            umma_bf16_2sm(d, ad0 + 6, bd0 + 6, idesc, 1u);
                        } else {
                            umma_bf16(d, ad0, bd0, idesc, acc_first);
                            umma_bf16(d, ad0 + 2, bd0 + 2, idesc, 1u);
                            umma_bf16(d, ad0 + 4, bd0 + 4, idesc, 1u);
                            umma_bf16(d, ad0 + 6, bd0 + 6, idesc, 1u);
                        }
                    }
                    if (PAIR2) {
                        umma_commit_2sm(b_empty(ib));
                        if (sg.a_last) umma_commit_2sm(a_empty(cur_a));
                    } else {
                        umma_commit(b_empty(ib));
                        if (sg.a_last) umma_commit(a_empty(cur_a));
                    }
                    if (++ib == (uint32_t)SB) { ib = 0; pb ^= 1; }
                }
                if (PAIR2) umma_commit_2sm(acc_full(as));
                else umma_commit(acc_full(as));
                CGN_STAMP(7);
            }
        }
    } else if (warp >= 2 + 8 * CGN_EPI_GROUPS) {
        // ===================== chain signalling (one thread) =====================
        // The epilogue warps only count themselves in (shared memory) once their stores of a sample have completed; this otherwise
        // idle thread then publishes the CTA's flag.  No gpu-scope release / atomic: a MEMBAR.GPU per sample cost 3 % of every
        // layer (it stalls behind all of the SM's stores in flight), and the flag is safe without it (see the producer warp).
        if (lane == 0 && F.done_flag != nullptr) {
            int it = 0;
            for (int b = grp; b < F.B; b += F.n_groups, ++it) {
                // two warpgroups on alternate samples: 8 warps arrive per sample of a group; ONE group: all 16 warps, every sample
                const uint32_t need = ONE_T ? 16u * (uint32_t)(it + 1) : 8u * (uint32_t)((it >> 1) + 1);
                const uint32_t sc = sig_cnt(ONE_T ? 0 : (it & 1));
                const long long t0 = clock64();
                for (;;) {
                    uint32_t v;
                    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(sc) : "memory");
                    if (v >= need) break;
                    __nanosleep(128);
                    if (clock64() - t0 > 4000000000LL) xchg_timeout(b, -2);
                }
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(F.done_flag + (size_t)b * CGN_MAX_G + j_cta), "r"(chain_v) : "memory");
            }
        }
    } else {
        // ===================== epilogue (warps 2..17) =====================
        constexpr bool ONE = ONE_T;
        constexpr int NT = ONE ? 512 : 256;           // threads of one epilogue group
        const int eg = ONE ? 0 : (warp - 2) >> 3;     // epilogue warpgroup: samples it = eg, eg + GROUPS, ... of this CTA
        const int e = ONE ? (warp - 2) : ((warp - 2) & 7);
        const int q = warp & 3;                       // TMEM lane quarter
        const int ch = (e >> 2) & 1;                  // which half of the columns
        const int kc_lo = ONE ? (e >> 3) : 0, kc_hi = ONE ? (e >> 3) + 1 : 2;     // my chunk(s) of the two per (quarter, half)
        const int tid = (threadIdx.x - 64) & (NT - 1);
        const int bar_id = 1 + eg;
        float* s_stat = s_grp0 + eg * GRP_FLOATS;
        float* s_x = s_stat + 256;
        float* s_abf = s_x + CGN_MAX_G * 16;
        float* s_gef = s_abf + 512;
        float* s_wf = s_gef + 512;
        const int cols_per_warp = P.bn >> 1;
        const uint32_t stg0 = sStage + (uint32_t)(warp - 2) * NBUF * 4096u;
        const uint32_t stgp0 = sPool + (uint32_t)(warp - 2) * NBUF * 2048u;
        float* my_stat = s_stat + e * 16;
        // this thread's coefficient column
        const bool has_col = tid < P.bn;
        const int cch = (n_tile * P.bn + (has_col ? tid : 0)) & (P.cout - 1);
        const int my_g = cch >> CG_LOG2;
        const int Cc = F.Cc;
        const float gw = F.gn_w[cch], gb = F.gn_b[cch], bs = bias ? bias[cch] : 0.0f;
        const float bcv = Cc > 0 ? F.bc[cch] : 0.0f;
        float wcj[NCA];
#pragma unroll
        for (int jj = 0; jj < NCA; ++jj) wcj[jj] = jj < Cc ? F.wc[cch * Cc + jj] : 0.0f;
        const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(F.ctrl) + 1u;
        const int step = F.step_ptr != nullptr ? *F.step_ptr : 0;
        const double inv_n = 1.0 / ((double)(1 << CG_LOG2) * (double)F.Lpos);
        // output phase (pair space) of this warp's columns: the CTA's n_tile, or the column half when a tile spans both
        const int my_par = !F.pair ? 0 : (P.n_tiles == 2 ? n_tile : ((ch * cols_per_warp) >= P.cout ? 1 : 0));
        int buf = 0, stores = 0;
        int b_sig = -1, k_groups = 0;                         // chaining: my previous sample (not yet signalled), store groups of the current one
        // Completion of a sample for the next layer: a warp whose store groups of its PREVIOUS sample have completed (deferred by
        // one sample, `keep` = groups of the newer sample still allowed in flight: no stall) arrives on the warpgroup's sig barrier;
        // the signalling warp publishes it (see above).
        auto stores_done = [&](int keep) {
            if (lane == 0) {
                if (keep >= 2) tma_wait_all<2>(); else if (keep == 1) tma_wait_all<1>(); else tma_wait_all<0>();
                asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(sig_cnt(eg)) : "memory");
            }
        };
        float f_gam = 0.0f, f_bet = 0.0f;                     // FiLM (gamma, beta) of my column for the current sample
        auto load_film = [&](int b) {
            const float* fr = F.film + (size_t)step * F.film_step_stride + (size_t)b * F.film_b_stride + F.film_off;
            f_gam = fr[cch];
            f_bet = fr[P.cout + cch];
        };
        if (has_col && grp + eg * F.n_groups < F.B) load_film(grp + eg * F.n_groups);
        auto load_cond = [&](int b, int mt, float (&c)[4][NCA]) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int row = row0 + mt * TC_BLOCK_M + q * 32 + 8 * m + (lane >> 2);
                const bool ok = row < P.rows;
                const int pos = F.pair ? 2 * row + my_par : row;
                const float* cp = F.cond + ((size_t)b * F.Lpos + pos) * Cc;
#pragma unroll
                for (int jj = 0; jj < NCA; ++jj) c[m][jj] = (ok && jj < Cc) ? cp[jj] : 0.0f;
            }
        };
        auto stage_wait = [&]() {
            if (stores >= NBUF) {
                if (lane == 0) {
                    if (NBUF == 2) tma_wait_read<1>(); else tma_wait_read<0>();
                }
                __syncwarp();
            }
        };
        constexpr int ITS = ONE ? 1 : CGN_EPI_GROUPS;
        for (int it = eg, b = grp + eg * F.n_groups; b < F.B; it += ITS, b += ITS * F.n_groups) {
            const int as = it & 1;
            if (tid == 0) CGN_STAMP(0);
            // ---- per-sample FiLM / cond coefficients of my column (nothing here depends on the statistics)
            if (has_col) {
                const float g = 1.0f + f_gam;
                const float ee = fmaf(bcv, g, f_bet);
                const int pi = tid >> 1, hf = tid & 1;
                s_gef[pi * 4 + hf] = g;
                s_gef[pi * 4 + 2 + hf] = ee;
#pragma unroll
                for (int jj = 0; jj < NCA; ++jj) s_wf[(pi * NCA + jj) * 2 + hf] = wcj[jj] * g;
                if (b + ITS * F.n_groups < F.B) load_film(b + ITS * F.n_groups);      // my next sample's row: latency hidden
            }
            if (lane < 16) my_stat[lane] = 0.0f;
            __syncwarp();
            // one warp polls the accumulator barrier, the other seven block on a hardware barrier: eight spinning warps were
            // ~20 % of all issued instructions, taken from the other warpgroup's arithmetic
            if (e == 0) mbar_wait(acc_full(as), ((uint32_t)(it >> 1)) & 1u);
            named_bar_sync(bar_id, NT);
            tc_fence_after();
            if (tid == 0) CGN_STAMP(1);
            // ---- pass 1: GroupNorm partial sums of (acc + bias) (+ the raw tensor for the backward pass)
#pragma unroll 1
            for (int kc = kc_lo; kc < kc_hi; ++kc) {
                const int mt = MT == 2 ? kc : 0;
                const int c0 = ch * cols_per_warp + (MT == 2 ? 0 : kc * 64);
                const int m_tile = ms * MT + mt;
                if (m_tile >= P.m_tiles) continue;
                const int row_base = m_tile * TC_BLOCK_M + q * 32;
                const uint32_t acc = tmem_base + (uint32_t)(as * 256 + mt * P.bn + c0) + ((uint32_t)(q * 32) << 16);
                const uint32_t stg = stg0 + buf * 4096;
                float sv[2 * (64 >> CG_LOG2)];
                if (F.write_raw) stage_wait();
                stat_chunk64<CG_LOG2>(acc, s_bias + c0, row_base + lane < P.rows, sv, F.write_raw != 0, stg, lane);
                if (F.write_raw) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(&tm_raw, stg, n_tile * P.bn + c0, row_base, b);
                        tma_commit();
                    }
                    ++stores;
                    ++k_groups;
                    if (NBUF == 2) buf ^= 1;
                }
                stat_reduce<CG_LOG2>(sv, my_stat, ((n_tile * P.bn + c0) & (P.cout - 1)) >> CG_LOG2, lane);
            }
            named_bar_sync(bar_id, NT);
            if (tid == 0) CGN_STAMP(2);
            // cond values of my fragment rows (row-group m: row 8m + lane/4 of this warp's 32 rows) in my output phase, for the
            // first pass-2 chunk: issued now, so the loads complete while the statistics travel
            float cd[4][NCA];
            load_cond(b, (ONE && MT == 2) ? kc_lo : 0, cd);
            // ---- exchange: publish this CTA's 16 sums, collect the G x 16 sums of the group
            if (tid < 16) {
                // fixed summation order, the same in both epilogue flavours (a sample's result must not depend on how many samples
                // the launch carries): per warp slot the two chunks first -- one warp adds them in the two-group flavour, warps w
                // and w + 8 hold them in the single-group flavour -- then the eight slots in order
                float v = 0.0f;
#pragma unroll
                for (int w = 0; w < 8; ++w) v += ONE ? s_stat[w * 16 + tid] + s_stat[(w + 8) * 16 + tid] : s_stat[w * 16 + tid];
                st_relaxed_u64(F.xchg + ((size_t)b * CGN_MAX_G + j_cta) * 16 + tid,
                               ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v));
            }
            for (int i = tid; i < G * 16; i += NT) {
                const unsigned long long* src = F.xchg + ((size_t)b * CGN_MAX_G + (i >> 4)) * 16 + (i & 15);
                unsigned long long pk = ld_relaxed_u64(src);
                if ((unsigned int)(pk >> 32) != epoch) {
                    const long long t0 = clock64();
                    do {
                        __nanosleep(64);                  // the chain is power-capped: do not burn issue slots and L2 requests while waiting
                        pk = ld_relaxed_u64(src);
                        if (clock64() - t0 > 4000000000LL) xchg_timeout(b, i >> 4);
                    } while ((unsigned int)(pk >> 32) != epoch);
                }
                s_x[i] = __uint_as_float((unsigned int)pk);
            }
            named_bar_sync(bar_id, NT);
            if (tid == 0) CGN_STAMP(3);
            // ---- statistics -> affine coefficients of my column (conv bias folded in; h = half of the GroupNorm output)
            if (has_col) {
                float f1 = 0.0f, f2 = 0.0f;
                for (int s = 0; s < G; ++s) {
                    f1 += s_x[s * 16 + my_g * 2];
                    f2 += s_x[s * 16 + my_g * 2 + 1];
                }
                const double mean = (double)f1 * inv_n;
                double var = (double)f2 * inv_n - mean * mean;
                if (var < 0.0) var = 0.0;
                const float rstd = rsqrtf((float)var + 1e-5f);
                const float meanf = (float)mean;
                const float a = rstd * gw;
                const float A2 = 0.5f * a;
                const float B2 = fmaf(A2, bs, 0.5f * (gb - meanf * a));
                const int pi = tid >> 1, hf = tid & 1;
                s_abf[pi * 4 + hf] = A2;
                s_abf[pi * 4 + 2 + hf] = B2;
                if (F.stats_out != nullptr && j_cta == 0 && tid < P.cout && (cch & ((1 << CG_LOG2) - 1)) == 0) {
                    F.stats_out[((size_t)b * 8 + my_g) * 2 + 0] = meanf;
                    F.stats_out[((size_t)b * 8 + my_g) * 2 + 1] = rstd;
                }
            }
            named_bar_sync(bar_id, NT);
            if (tid == 0) CGN_STAMP(4);
            // ---- pass 2: normalise / activate / modulate out of TMEM
#pragma unroll 1
            for (int kc = kc_lo; kc < kc_hi; ++kc) {
                const int mt = MT == 2 ? kc : 0;
                const int c0 = ch * cols_per_warp + (MT == 2 ? 0 : kc * 64);
                const int m_tile = ms * MT + mt;
                if (m_tile >= P.m_tiles) continue;
                const int row_base = m_tile * TC_BLOCK_M + q * 32;
                const uint32_t acc = tmem_base + (uint32_t)(as * 256 + mt * P.bn + c0) + ((uint32_t)(q * 32) << 16);
                const uint32_t stg = stg0 + buf * 4096, stgp = stgp0 + buf * 2048;
                float cdn[4][NCA];
                if (!ONE && MT == 2 && kc == 0) load_cond(b, 1, cdn);      // the second tile's rows, needed one chunk later
                stage_wait();
                const ulonglong2* ab = reinterpret_cast<const ulonglong2*>(s_abf) + (c0 >> 1);
                const ulonglong2* ge = reinterpret_cast<const ulonglong2*>(s_gef) + (c0 >> 1);
                const unsigned long long* wv = reinterpret_cast<const unsigned long long*>(s_wf) + (c0 >> 1) * NCA;
                unsigned long long hd[4][3];
                if (HEAD) {
#pragma unroll
                    for (int m = 0; m < 4; ++m) hd[m][0] = hd[m][1] = hd[m][2] = 0ull;
                }
                const bool st_out = !HEAD || F.store_out != 0;
                if (Cc > 0) {
#pragma unroll 1
                    for (int kb = 0; kb < 4; ++kb)
                        gn_cols16<NCA, true, POOL, HEAD>(acc, kb, ab, ge, wv, cd, stg, stgp, lane, F.dbg_mode, s_hw, hd, st_out);
                } else {
#pragma unroll 1
                    for (int kb = 0; kb < 4; ++kb)
                        gn_cols16<NCA, false, POOL, HEAD>(acc, kb, ab, ge, wv, cd, stg, stgp, lane, F.dbg_mode, s_hw, hd, st_out);
                }
                if (HEAD) {
                    // fold the channel pair, then the 4 lanes that share a row; lane (tr, tq) writes row-group m = tq
                    const int tq = lane & 3, trow = lane >> 2;
                    float dsel[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
#pragma unroll
                        for (int tp = 0; tp < 3; ++tp) {
                            float lo, hi;
                            upk2(hd[m][tp], lo, hi);
                            float v = lo + hi;
                            v += __shfl_xor_sync(0xffffffffu, v, 1);
                            v += __shfl_xor_sync(0xffffffffu, v, 2);
                            if (tq == m) dsel[tp] = v;
                        }
                    }
                    const int row = row_base + trow + 8 * tq;
                    if (row < P.rows) {
                        const int pos = F.pair ? 2 * row + my_par : row;
                        *reinterpret_cast<float4*>(F.head_dots + ((size_t)b * F.Lpos + pos) * 4) =
                            make_float4(dsel[0], dsel[1], dsel[2], 0.0f);
                    }
                    if (!st_out) {
                        if (!ONE && MT == 2 && kc == 0) {
#pragma unroll
                            for (int m = 0; m < 4; ++m) {
#pragma unroll
                                for (int jj = 0; jj < NCA; ++jj) cd[m][jj] = cdn[m][jj];
                            }
                        }
                        continue;
                    }
                }
                fence_proxy_async();
                __syncwarp();
#ifdef CGN_ABLATE
                if (F.dbg_mode & 8) continue;
#endif
                if (lane == 0) {
#ifdef CGN_ABLATE
                    if (!(F.dbg_mode & 128))
#endif
                    tma_store_3d(&tm_out, stg, n_tile * P.bn + c0, row_base, b);
#ifdef CGN_ABLATE
                    if (!(F.dbg_mode & 64))
#endif
                    if (POOL) tma_store_3d(&tm_pool, stgp, n_tile * P.bn + c0, row_base >> 1, b);
                    tma_commit();
                }
                ++stores;
                ++k_groups;
                if (NBUF == 2) buf ^= 1;
                if (!ONE && MT == 2 && kc == 0) {
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
#pragma unroll
                        for (int jj = 0; jj < NCA; ++jj) cd[m][jj] = cdn[m][jj];
                    }
                }
            }
            // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
            tc_fence_before();
            named_bar_sync(bar_id, NT);
            if (e == 0 && lane == 0) {
                if (leader) mbar_arrive(acc_empty(as));
                else mbar_arrive_cluster(mapa_u32(acc_empty(as), 0));       // the leader issues this pair's MMAs
            }
            if (F.done_flag != nullptr) {
                // after the accumulator stage went back to the MMA warp (off its critical path): my stores of the PREVIOUS sample
                if (b_sig >= 0) stores_done(k_groups);
                b_sig = b;
                k_groups = 0;
            }
            if (tid == 0) CGN_STAMP(5);
        }
        if (F.done_flag != nullptr && b_sig >= 0) stores_done(0);
        if (lane == 0) tma_wait_all<0>();
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR2) cluster_sync_all();          // neither CTA leaves (or frees its TMEM) while the pair's MMAs / remote arrives are in flight
    if (warp == 2) {
        tc_fence_after();
        if (PAIR2) tmem_dealloc_2sm(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
    // last CTA out advances the epoch for the next launch (every CTA read it before it could finish)
    if (threadIdx.x == 0) {
        if (F.tdbg != nullptr) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            F.tdbg[blockIdx.x * 2 + 1] = t;
        }
        __threadfence();
        const unsigned int done = atomicAdd(F.ctrl + 1, 1u);
        if (done == gridDim.x - 1) {
            F.ctrl[1] = 0u;
            __threadfence();
            atomicAdd(F.ctrl, 1u);
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct CgnPlan {
    TcParams P;
    int MT, G, n_groups, smem, lg;
    int pair2, wbox;          // CTA pairs (cta_group::2) and the weight TMA box rows they need
};
// gw_set_option("pair2", 0/1).  Measured on B200 (profiles/r02_pair2.md): the pair flavour -- half the weight bytes per SM, B ring
// twice as deep -- leaves every layer's MMA-issue time unchanged (dec0: 20.0 k vs 20.3 k cycles per sample slice), i.e. the
// operand supply is not what paces the MMA warp; it stays a parity-tested option, off by default.
int g_cgn_pair2 = 0;
// gw_set_option("one_group", 0/1): allow the single-group epilogue flavour for launches with at most one sample per CTA group
int g_cgn_one_group = 1;

static int cgn_plan(const gw_conv_tc_shape* s, int Cc, bool pool, CgnPlan* pl) {
    int rc = build_params(s, &pl->P, true);
    if (rc != GW_OK) return rc;
    TcParams& P = pl->P;
    GW_REQUIRE(s->pair == 0 || s->pair == 1, "conv_gn: pair mode %d is a dgrad mode", s->pair);
    GW_REQUIRE(P.bn == 128 || P.bn == 256, "conv_gn: accumulator tile of %d columns (need 128 or 256)", P.bn);
    GW_REQUIRE(s->Cout == 64 || s->Cout == 128 || s->Cout == 256, "conv_gn: Cout=%d", s->Cout);
    GW_REQUIRE(!pool || (s->pair == 0 && s->L % 2 == 0), "conv_gn: pooling needs a plain conv of even length");
    GW_REQUIRE(Cc >= 0 && Cc <= CGN_NCA_MAX, "conv_gn: Cc=%d", Cc);
    pl->MT = 256 / P.bn;
    pl->lg = s->Cout == 64 ? 3 : (s->Cout == 128 ? 4 : 5);
    const bool combo = (pl->lg == 3 && pl->MT == 2 && s->pair == 1) || (pl->lg == 4 && pl->MT == 2 && s->pair == 0) ||
                       (pl->lg == 4 && pl->MT == 1 && s->pair == 1) || (pl->lg == 5 && pl->MT == 1);
    GW_REQUIRE(combo, "conv_gn: unsupported layer shape (Cout=%d pair=%d)", s->Cout, s->pair);
    GW_REQUIRE(P.rows % 32 == 0, "conv_gn: rows=%d must be a multiple of 32", P.rows);
    pl->G = gw_cdiv(P.m_tiles, pl->MT) * P.n_tiles;
    GW_REQUIRE(pl->G <= CGN_MAX_G, "conv_gn: a sample needs %d CTAs (max %d)", pl->G, CGN_MAX_G);
    const int sms = sm_count_cached();
    GW_REQUIRE(pl->G <= sms, "conv_gn: group larger than the GPU");
    pl->n_groups = sms / pl->G;
    if (pl->n_groups > s->B) pl->n_groups = s->B;
    const int nca = (Cc == 1 || Cc == 5) ? Cc : CGN_NCA_MAX;
    const int misc = 256 + 1024 + CGN_EPI_GROUPS * (1024 + CGN_MAX_G * 64 + 2048 + 2048 + 1024 * nca) + 64 + (pl->lg == 3 ? 768 : 0);
    // CTA pairs: the row slices of a sample must pair up, and half of every weight tile must be a whole number of 32-row boxes
    const int m_slices = pl->G / P.n_tiles;
    pl->pair2 = g_cgn_pair2 && (m_slices % 2 == 0);
    pl->wbox = 64;
    for (int t = 0; t < P.n_tiles && pl->pair2; ++t)
        for (int i = 0; i < P.n_seg[t]; ++i) {
            const int half = P.seg[t][i].n_cnt / 2;
            if (half % 32 != 0) pl->pair2 = 0;
            else if (half % 64 != 0) pl->wbox = 32;
        }
    if (!pl->pair2) pl->wbox = 64;
    // ring / staging depths: prefer deep rings, shrink until the CTA fits.  The MMA warp is paced by the operand rings: a ring
    // holds (slots x bytes) in flight against ~1-1.5 us of L2 -> shared-memory latency, so pairs (half-size B slots) go deeper
    const int ew = 8 * CGN_EPI_GROUPS;
    const int cand1[6][3] = {{2, 4, 2}, {2, 3, 2}, {2, 4, 1}, {2, 3, 1}, {2, 2, 1}, {1, 2, 1}};
    const int cand2[8][3] = {{3, 8, 2}, {3, 6, 2}, {3, 8, 1}, {3, 7, 1}, {3, 6, 1}, {2, 6, 1}, {2, 5, 1}, {2, 4, 1}};
    const int n_cand = pl->pair2 ? 8 : 6;
    const int b_slot = P.bn * (pl->pair2 ? 64 : 128);
    pl->smem = 0;
    for (int i = 0; i < n_cand; ++i) {
        const int sa = pl->pair2 ? cand2[i][0] : cand1[i][0], sb = pl->pair2 ? cand2[i][1] : cand1[i][1];
        const int nb = pl->pair2 ? cand2[i][2] : cand1[i][2];
        const int sm = 1024 + sa * pl->MT * TC2_A_SLOT + sb * b_slot + ew * nb * 4096 + (pool ? ew * nb * 2048 : 0) + misc;
        if (sm <= 232448) {
            P.sa = sa; P.sb = sb; P.nbuf = nb; P.n_acc = 2;
            pl->smem = sm;
            break;
        }
    }
    GW_REQUIRE(pl->smem > 0, "conv_gn: shared memory does not fit");
    return GW_OK;
}

// number of CTAs that share one sample (0 if the fused kernel cannot run this layer); see gw_conv_gn
extern "C" int gw_conv_gn_group(const gw_conv_tc_shape* s, int Cc, int pool) {
    CgnPlan pl;
    if (cgn_plan(s, Cc, pool != 0, &pl) != GW_OK) return 0;
    return pl.G;
}
static long long* g_cgn_dbg = nullptr;
static unsigned long long* g_cgn_tdbg = nullptr;      // tools/chain_overlap.py: [launch][160 CTAs][2] globaltimer stamps
static int g_cgn_tdbg_launch = 0;
extern "C" void gw_conv_gn_tdebug(void* buf) { g_cgn_tdbg = (unsigned long long*)buf; g_cgn_tdbg_launch = 0; }
static int g_cgn_dbg_mode = 0;
extern "C" void gw_conv_gn_debug(void* buf) { g_cgn_dbg = (long long*)buf; }
extern "C" void gw_conv_gn_debug_mode(int m) { g_cgn_dbg_mode = m; }     // tools/cgn_timeline.py
// [0, 64): epoch / finished counter; packets [B][CGN_MAX_G][16] x 8 bytes; chain flags [B][CGN_MAX_G] (uint32)
extern "C" long gw_conv_gn_sync_bytes(int B) { return 64 + (long)B * CGN_MAX_G * 16 * 8 + (long)B * CGN_MAX_G * 4 + 64; }

// gw_conv_gn2 = gw_conv_gn + the head conv's dot products (last decoder, Cout = 64 in pair space): head_w = final.weight
// [C+1, 3], head_dots [B, L, 4] fp32 receive (sum_c out[l,c] w[c,k])_k for gw_final_step(dtype = GW_DOTS); out == NULL then
// skips the activated tensor altogether.
// gw_conv_gn3 = gw_conv_gn2 + layer chaining for the sampler: serial_ptr != NULL (with step_ptr) makes this launch signal per-sample
// completion in its sync buffer; prev_sync = the sync buffer of the layer that produced src0 in the SAME reverse step (launched
// with a non-NULL serial_ptr): this launch then overlaps the tail of that kernel and orders itself per sample.
extern "C" int gw_conv_gn3(const gw_conv_tc_shape* s, const void* src0, const void* src1, const void* packed, const float* bias,
                           const float* gn_w, const float* gn_b, const float* cond, int Cc, const float* wc, const float* bc,
                           const float* film, int film_off, long film_b_stride, long film_step_stride, const int* step_ptr,
                           void* out, void* pooled, void* raw, float* stats_out, void* sync_buf, const float* head_w,
                           float* head_dots, const void* prev_sync, int prev_G, const int* serial_ptr, void* stream) {
    CgnPlan pl;
    int rc = cgn_plan(s, Cc, pooled != nullptr, &pl);
    if (rc != GW_OK) return rc;
    TcParams& P = pl.P;
    const bool head = head_dots != nullptr;
    GW_REQUIRE(src0 != nullptr && packed != nullptr && (out != nullptr || head) && sync_buf != nullptr && gn_w != nullptr &&
                   gn_b != nullptr && film != nullptr,
               "conv_gn: null pointer");
    GW_REQUIRE(!head || (head_w != nullptr && pl.lg == 3 && pooled == nullptr), "conv_gn: the head dots need Cout = 64 in pair space");
    GW_REQUIRE((s->n_src == 2) == (src1 != nullptr), "conv_gn: src1 / n_src mismatch");
    GW_REQUIRE((cond != nullptr) == (Cc > 0) && (Cc == 0 || (wc != nullptr && bc != nullptr)), "conv_gn: cond/Cc mismatch");
    CUtensorMap ta0, ta1, tw, to, tr, tp;
    if ((rc = make_map3(&ta0, src0, (uint64_t)s->C0, (uint64_t)s->L0, (uint64_t)s->B, 64, 130)) != GW_OK) return rc;
    const uint64_t oc = s->pair == 1 ? 2 * (uint64_t)s->Cout : (uint64_t)s->Cout;
    if (s->pair == 1) {
        if ((rc = make_map3(&ta1, src1, (uint64_t)2 * s->C1, (uint64_t)P.rows, (uint64_t)s->B, 64, 130)) != GW_OK) return rc;
    } else {
        ta1 = ta0;
    }
    // (out == NULL with head dots: the map is never used; any valid bf16 tensor of the right extent will do -> src1 / src0)
    const void* out_map = out != nullptr ? out : (raw != nullptr ? raw : nullptr);
    if (out_map != nullptr) {
        if ((rc = make_map3(&to, out_map, oc, (uint64_t)P.rows, (uint64_t)s->B, 64, 32)) != GW_OK) return rc;
    } else {
        to = ta0;
    }
    tr = to;
    if (raw != nullptr && (rc = make_map3(&tr, raw, oc, (uint64_t)P.rows, (uint64_t)s->B, 64, 32)) != GW_OK) return rc;
    tp = to;
    if (pooled != nullptr && (rc = make_map3(&tp, pooled, oc, (uint64_t)(P.rows / 2), (uint64_t)s->B, 64, 16)) != GW_OK) return rc;
    if ((rc = make_map2(&tw, packed, (uint64_t)P.k_total, (uint64_t)P.n_tiles * P.bn, 64, (uint32_t)pl.wbox)) != GW_OK) return rc;
    GnFuseArgs F;
    F.gn_w = gn_w; F.gn_b = gn_b; F.cond = cond; F.wc = wc; F.bc = bc; F.film = film; F.step_ptr = step_ptr;
    F.stats_out = stats_out;
    F.ctrl = reinterpret_cast<unsigned int*>(sync_buf);
    F.xchg = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(sync_buf) + 64);
    F.film_b_stride = film_b_stride; F.film_step_stride = film_step_stride; F.film_off = film_off;
    F.Cc = Cc; F.G = pl.G; F.n_groups = pl.n_groups; F.B = s->B; F.Lpos = s->L; F.write_raw = raw != nullptr ? 1 : 0;
    F.pair = s->pair == 1 ? 1 : 0;
    F.head_w = head_w; F.head_dots = head_dots; F.store_out = out != nullptr ? 1 : 0;
    F.pair2 = pl.pair2; F.wbox = pl.wbox;
    {
        const size_t flags_off = 64 + (size_t)s->B * CGN_MAX_G * 16 * 8;
        const bool sig = serial_ptr != nullptr && step_ptr != nullptr;
        GW_REQUIRE(prev_sync == nullptr || (sig && prev_G >= 1 && prev_G <= 8), "conv_gn: chaining needs serial_ptr, step_ptr and a producer group of <= 8 CTAs");
        F.serial_ptr = sig ? serial_ptr : nullptr;
        F.done_flag = sig ? reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(sync_buf) + flags_off) : nullptr;
        F.prev_flag = prev_sync != nullptr ? reinterpret_cast<const unsigned int*>(reinterpret_cast<const uint8_t*>(prev_sync) + flags_off)
                                           : nullptr;
        F.prev_G = prev_G;
    }
    const bool chained = F.prev_flag != nullptr;
    F.dbg = g_cgn_dbg;
    F.tdbg = g_cgn_tdbg != nullptr ? g_cgn_tdbg + (size_t)(g_cgn_tdbg_launch++) * 160 * 2 : nullptr;
    F.dbg_mode = g_cgn_dbg_mode;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = pl.G * pl.n_groups, smem = pl.smem;
    const bool pool = pooled != nullptr;
    // one sample per CTA group at most: the single-group epilogue flavour (all 16 warps on that sample)
    const bool one = !pl.pair2 && g_cgn_one_group != 0 && s->B <= pl.n_groups;
#define CGN_GO3(LG, MTV, CCV, PL, HD, P2, ON)                                                                             \
    do {                                                                                                                  \
        GW_CUDA(cudaFuncSetAttribute(conv_gn_kernel<LG, MTV, CCV, PL, HD, P2, ON>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        GW_CUDA(gw_launch_cluster(conv_gn_kernel<LG, MTV, CCV, PL, HD, P2, ON>, dim3(grid), dim3(CGN_THREADS), (size_t)smem, st, P2 ? 2 : 1, chained, ta0, ta1, tw, to, tr, tp, P, F, bias)); \
    } while (0)
#define CGN_GO2(LG, MTV, CCV, PL, HD, P2)                                                                                 \
    do {                                                                                                                  \
        if (!(P2) && one) CGN_GO3(LG, MTV, CCV, PL, HD, false, true);                                                     \
        else CGN_GO3(LG, MTV, CCV, PL, HD, P2, (CGN_ONE_GROUP != 0));                                                     \
    } while (0)
#define CGN_GO(LG, MTV, CCV, PL)                                  \
    do {                                                          \
        if (pl.pair2) CGN_GO2(LG, MTV, CCV, PL, false, true);     \
        else CGN_GO2(LG, MTV, CCV, PL, false, false);             \
    } while (0)
#define CGN_CC(LG, MTV, PL)                      \
    do {                                         \
        if (Cc == 1) CGN_GO(LG, MTV, 1, PL);     \
        else if (Cc == 5) CGN_GO(LG, MTV, 5, PL);\
        else CGN_GO(LG, MTV, -1, PL);            \
    } while (0)
#define CGN_GOH(CCV)                                              \
    do {                                                          \
        if (pl.pair2) CGN_GO2(3, 2, CCV, false, true, true);      \
        else CGN_GO2(3, 2, CCV, false, true, false);              \
    } while (0)
    if (pl.lg == 3 && head) {
        if (Cc == 1) CGN_GOH(1);
        else if (Cc == 5) CGN_GOH(5);
        else CGN_GOH(-1);
    } else if (pl.lg == 3) CGN_CC(3, 2, false);
    else if (pl.lg == 4 && pl.MT == 2) { if (pool) CGN_CC(4, 2, true); else CGN_CC(4, 2, false); }
    else if (pl.lg == 4) CGN_CC(4, 1, false);
    else { if (pool) CGN_CC(5, 1, true); else CGN_CC(5, 1, false); }
#undef CGN_GOH
#undef CGN_CC
#undef CGN_GO
#undef CGN_GO2
#undef CGN_GO3
    GW_LAUNCH_CHECK();
    return GW_OK;
}
extern "C" int gw_conv_gn2(const gw_conv_tc_shape* s, const void* src0, const void* src1, const void* packed, const float* bias,
                           const float* gn_w, const float* gn_b, const float* cond, int Cc, const float* wc, const float* bc,
                           const float* film, int film_off, long film_b_stride, long film_step_stride, const int* step_ptr,
                           void* out, void* pooled, void* raw, float* stats_out, void* sync_buf, const float* head_w,
                           float* head_dots, void* stream) {
    return gw_conv_gn3(s, src0, src1, packed, bias, gn_w, gn_b, cond, Cc, wc, bc, film, film_off, film_b_stride, film_step_stride,
                       step_ptr, out, pooled, raw, stats_out, sync_buf, head_w, head_dots, nullptr, 0, nullptr, stream);
}
extern "C" int gw_conv_gn(const gw_conv_tc_shape* s, const void* src0, const void* src1, const void* packed, const float* bias,
                          const float* gn_w, const float* gn_b, const float* cond, int Cc, const float* wc, const float* bc,
                          const float* film, int film_off, long film_b_stride, long film_step_stride, const int* step_ptr,
                          void* out, void* pooled, void* raw, float* stats_out, void* sync_buf, void* stream) {
    return gw_conv_gn2(s, src0, src1, packed, bias, gn_w, gn_b, cond, Cc, wc, bc, film, film_off, film_b_stride, film_step_stride,
                       step_ptr, out, pooled, raw, stats_out, sync_buf, nullptr, nullptr, stream);
}
