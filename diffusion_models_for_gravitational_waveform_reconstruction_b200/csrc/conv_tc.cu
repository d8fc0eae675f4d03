// tcgen05 / TMA implicit-GEMM Conv1d(k=3, pad=1) on bf16 channels-last activations (sm_100a).
//
// GEMM view:  D[row, n] = sum over "segments" A_seg[row + shift, 64 channels] * W_seg[n, 64 channels]^T
//   * rows are positions of one sample (M tile = 128 rows), n are output channels;
//   * a segment is one 64-channel K chunk of one source tensor at one row shift (-1, 0, +1): the three conv taps
//     are three shifted TMA loads of the same channels-last tensor, and the zero padding of the conv is the TMA
//     out-of-bounds zero fill (3-D tensor map [C, rows, B] so a halo never crosses into the neighbouring sample);
//   * decoder convs run in "pair space": the input is cat[nearest-upsample(h), skip] (models.py:217-222).  With
//     row m standing for output positions (2m, 2m+1) and N = (phase, cout), the upsampled half needs only h[m-1],
//     h[m], h[m+1] with tap-summed weights (W1+W2 | W0+W1), and the skip half is the same memory viewed as
//     [B, L/2, 2*C1].  No upsampled tensor and no concat is ever materialised, and the upsampled half costs 4 instead
//     of 6 MMA blocks.
// Pipeline: warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma issuer (accumulator in TMEM),
// warps 2-5 = epilogue (tcgen05.ld -> +bias -> bf16 -> GroupNorm partial sums -> swizzled smem -> TMA store).
#include "common.cuh"
#include "../../include/gwb200.h"
#include "tc_common.cuh"
#include <string.h>

#define TC_MAX_SEG 48
#define TC_BLOCK_M 128
#define TC_BLOCK_K 64

struct TcSeg {
    int16_t src;      // 0: src0, 1: src1 (pair view)
    int16_t shift;    // row shift -1/0/+1
    int16_t col;      // first channel (column of the source view)
    int16_t n_off;    // first accumulator column this segment updates
    int16_t n_cnt;    // number of accumulator columns
    int16_t ci0;      // first reference input channel of this chunk (packing only)
    uint8_t mask[2];  // taps summed into the weights of phase 0 / phase 1 (packing only)
    int32_t wk;       // K offset of this segment in the packed weight matrix
    // v2 kernel: segments that read the same 64-channel column share one A tile of 130 rows (row0-1 .. row0+128)
    uint8_t a_new;    // load a new A tile before this segment
    uint8_t a_last;   // release the A tile after this segment
    int8_t load_shift;  // row shift of that TMA load
    int8_t desc_row;    // row (0..2) inside the A tile where this segment's MMA operand starts
};

struct TcParams {
    int n_seg[2];
    TcSeg seg[2][TC_MAX_SEG];
    int rows;      // rows per sample in row space (L, or L/2 in pair space)
    int m_tiles;   // ceil(rows / 128)
    int n_tiles;   // 1 or 2
    int bn;        // accumulator columns per tile
    int cout;      // channels of the conv output
    int stages;
    int k_total;   // packed K extent (max n_seg * 64)
    // v2 kernel
    int m_super;   // ceil(m_tiles / 2)
    int total_tiles;
    int sa, sb, nbuf, n_acc;   // A ring slots, B ring slots, store staging buffers per warp, accumulator stages
    int use_base_off;
};

// ------------------------------------------------------------------------------------------------ epilogue

// One epilogue unit: 64 accumulator columns of this lane's row -> +bias -> bf16 -> 128-byte-swizzled staging rows, plus
// (optionally) the GroupNorm partial sums of the fp32 values.  CG_LOG2 = log2(channels per group) in {3, 4, 5}.
// stat: this warp's [8 groups][2] slots (smem) for the current M tile; grp0: group of the chunk's first column.
template <int CG_LOG2, bool STORE = true>
__device__ __forceinline__ void epi_chunk64(uint32_t taddr, const float* sbias, bool row_ok, bool do_stats, float* stat, int grp0,
                                            uint32_t stg, int lane) {
    constexpr int NG = 64 >> CG_LOG2;                         // groups inside the 64 columns: 8, 4, 2
    constexpr int PG = (1 << CG_LOG2) / 2;                    // column PAIRS per group
    unsigned long long a1[NG], a2[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) { a1[g] = 0ull; a2[g] = 0ull; }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(hh * 32), v);
        unsigned long long x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(sbias + hh * 32 + 2 * i);
            x[i] = fadd2(pk2(v[2 * i], v[2 * i + 1]), bb);
            const int g = (hh * 16 + i) / PG;
            a1[g] = fadd2(a1[g], x[i]);
            a2[g] = ffma2(x[i], x[i], a2[g]);
        }
        // 32 columns = 4 chunks of 16 B; SW128: chunk c of row r lives at chunk (c ^ (r & 7))
#pragma unroll
        for (int j = 0; j < (STORE ? 4 : 0); ++j) {
            uint32_t pkd[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float lo, hi;
                upk2(x[j * 4 + q], lo, hi);
                pkd[q] = pack_bf16x2(lo, hi);
            }
            const int chunk = (hh * 4 + j) ^ (lane & 7);
            const uint32_t dst = stg + (uint32_t)lane * 128 + (uint32_t)chunk * 16;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pkd[0]), "r"(pkd[1]), "r"(pkd[2]), "r"(pkd[3])
                         : "memory");
        }
    }
    if (do_stats) {
        float sv[2 * NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            float lo, hi;
            upk2(a1[g], lo, hi);
            sv[2 * g] = row_ok ? lo + hi : 0.0f;
            upk2(a2[g], lo, hi);
            sv[2 * g + 1] = row_ok ? lo + hi : 0.0f;
        }
        const float tot = warp_reduce_multi<2 * NG>(sv, lane);
        constexpr int LOW = 2 * NG == 16 ? 2 : (2 * NG == 8 ? 4 : 8);   // lanes sharing one value
        if ((lane & (LOW - 1)) == 0) stat[grp0 * 2 + lane / LOW] += tot;
    }
}

// ------------------------------------------------------------------------------------------------ kernel
// CG_LOG2: log2(channels per GroupNorm group) clipped to 5 (a 32-column chunk then lies inside one group)
template <int CG_LOG2>
__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
               const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_out,
               const __grid_constant__ TcParams P, const float* __restrict__ bias, float* __restrict__ part) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int stages = P.stages;
    const uint32_t a_bytes = TC_BLOCK_M * TC_BLOCK_K * 2;          // 16 KB
    const uint32_t b_bytes = (uint32_t)P.bn * TC_BLOCK_K * 2;      // up to 32 KB
    const uint32_t sA = base;
    const uint32_t sB = sA + stages * a_bytes;
    const uint32_t sStage = sB + stages * b_bytes;                 // 4 warps x 2 buffers x 4 KB
    const uint32_t sMisc = sStage + 4 * 2 * 4096;
    // misc region (generic pointers)
    uint8_t* misc = smem_raw + (sMisc - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(misc);            // full[stages], empty[stages], tmem_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 8 * 20);
    float* s_bias = reinterpret_cast<float*>(misc + 256);          // [bn]
    float* s_stat = s_bias + 256;                                  // [4 warps][8 groups][2]
    auto full_bar = [&](int s) { return smem_u32(bars + s); };
    auto empty_bar = [&](int s) { return smem_u32(bars + stages + s); };
    const uint32_t tmem_full_bar = smem_u32(bars + 2 * stages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x % P.m_tiles, b = blockIdx.x / P.m_tiles, n_tile = blockIdx.y;
    const int row0 = m_tile * TC_BLOCK_M;
    const int n_seg = P.n_seg[n_tile];
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < P.bn) tmem_cols <<= 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_a0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_a1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_out) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    for (int i = threadIdx.x; i < P.bn; i += blockDim.x) s_bias[i] = bias ? bias[(n_tile * P.bn + i) % P.cout] : 0.0f;
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_stat[i] = 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int s = 0; s < n_seg; ++s) {
                const TcSeg sg = P.seg[n_tile][s];
                mbar_wait(empty_bar(stage), phase ^ 1);
                mbar_expect_tx(full_bar(stage), a_bytes + (uint32_t)sg.n_cnt * TC_BLOCK_K * 2);
                tma_load_3d(sA + stage * a_bytes, sg.src ? &tm_a1 : &tm_a0, full_bar(stage), sg.col, row0 + sg.shift, b);
                const int wrow = n_tile * P.bn + sg.n_off;
                for (int j = 0; j < sg.n_cnt; j += 64)
                    tma_load_2d(sB + stage * b_bytes + (uint32_t)j * 128, &tm_w, full_bar(stage), sg.wk, wrow + j);
                if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int s = 0; s < n_seg; ++s) {
                const TcSeg sg = P.seg[n_tile][s];
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t idesc = make_idesc(TC_BLOCK_M, (uint32_t)sg.n_cnt);
                const uint32_t a0 = sA + stage * a_bytes, b0 = sB + stage * b_bytes;
#pragma unroll
                for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
                    const uint64_t ad = make_sw128_desc(a0 + k * 32, 0);
                    const uint64_t bd = make_sw128_desc(b0 + k * 32, 0);
                    umma_bf16(tmem_base + (uint32_t)sg.n_off, ad, bd, idesc, (s > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(empty_bar(stage));
                if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(tmem_full_bar);
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                       // TMEM lane quarter this warp may touch
        const int row = row0 + q * 32 + lane;
        const bool row_ok = row < P.rows;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const uint32_t stg0 = sStage + (uint32_t)(warp - 2) * 2 * 4096;
        float* my_stat = s_stat + (warp - 2) * 16;
        const int cg = P.cout >> 3;
        int buf = 0;
        for (int c0 = 0; c0 < P.bn; c0 += 64) {
            const uint32_t stg = stg0 + buf * 4096;
            if (c0 >= 128) {                          // the buffer we are about to overwrite was stored two blocks ago
                if (lane == 0) tma_wait_read<1>();
                __syncwarp();
            }
            {
                const int ch0 = (n_tile * P.bn + c0) & (P.cout - 1);
                epi_chunk64<CG_LOG2>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, s_bias + c0, row_ok, part != nullptr,
                                     my_stat, ch0 >> CG_LOG2, stg, lane);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_store_3d(&tm_out, stg, n_tile * P.bn + c0, row0 + q * 32, b);
                tma_commit();
            }
            buf ^= 1;
        }
        if (lane == 0) tma_wait_all<0>();
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
    if (threadIdx.x < 8 && part != nullptr) {
        const int g = threadIdx.x;
        float a1 = 0.0f, a2 = 0.0f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            a1 += s_stat[w * 16 + g * 2 + 0];
            a2 += s_stat[w * 16 + g * 2 + 1];
        }
        const int n_part = P.m_tiles * P.n_tiles;
        float* pt = part + ((size_t)b * n_part + (m_tile * P.n_tiles + n_tile)) * 16;
        pt[g * 2 + 0] = a1;
        pt[g * 2 + 1] = a2;
    }
}

// ------------------------------------------------------------------------------------------------ kernel v2
// Persistent CTAs (one per SM).  A work item is a SUPER-TILE of 2 x 128 rows x bn columns of one sample: every weight
// tile fetched from L2 feeds two accumulators (the L2->SM fabric, ~42 B/clk/SM, is what bounds the v1 kernel), and the
// three taps of a 64-channel chunk read ONE 130-row A tile through row-shifted UMMA descriptors (halo mode).
// warp 0: TMA producer | warp 1: MMA issuer | warps 2..9: epilogue (two warps per TMEM lane quarter, half the columns each).
#define TC2_A_SLOT 17408          // 130 rows x 128 B rounded up to 1 KB
#define TC2_A_BYTES 16640         // bytes one A box really transfers
#define TC2_MT 2

template <int CG_LOG2>
__global__ void __launch_bounds__(320, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
                const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_out,
                const __grid_constant__ TcParams P, const float* __restrict__ bias, float* __restrict__ part) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int SA = P.sa, SB = P.sb, NBUF = P.nbuf, NACC = P.n_acc;
    const uint32_t b_bytes = (uint32_t)P.bn * TC_BLOCK_K * 2;
    const uint32_t sA = base;                                           // [SA][MT][A_SLOT]
    const uint32_t sB = sA + (uint32_t)SA * TC2_MT * TC2_A_SLOT;        // [SB][b_bytes]
    const uint32_t sStage = sB + (uint32_t)SB * b_bytes;                // [8 warps][NBUF][4096]
    const uint32_t sMisc = sStage + 8u * NBUF * 4096u;
    uint8_t* misc = smem_raw + (sMisc - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(misc);                 // a_full[SA] a_empty[SA] b_full[SB] b_empty[SB] acc_full[2] acc_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 8 * 28);
    float* s_bias = reinterpret_cast<float*>(misc + 256);               // [2][bn]  (per n_tile)
    float* s_stat = s_bias + 512;                                       // [2 parity][MT][8 warps][8 groups][2]
    auto a_full = [&](int i) { return smem_u32(bars + i); };
    auto a_empty = [&](int i) { return smem_u32(bars + SA + i); };
    auto b_full = [&](int i) { return smem_u32(bars + 2 * SA + i); };
    auto b_empty = [&](int i) { return smem_u32(bars + 2 * SA + SB + i); };
    auto acc_full = [&](int i) { return smem_u32(bars + 2 * SA + 2 * SB + i); };
    auto acc_empty = [&](int i) { return smem_u32(bars + 2 * SA + 2 * SB + 2 + i); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < NACC * TC2_MT * P.bn) tmem_cols <<= 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_a0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_a1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_out) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < SA; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
        for (int i = 0; i < SB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), 1); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    for (int i = threadIdx.x; i < P.n_tiles * P.bn; i += blockDim.x) s_bias[i] = bias ? bias[i % P.cout] : 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t ia = 0, pa = 0, ib = 0, pb = 0;
            for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
                const int n_tile = tile % P.n_tiles;
                const int r = tile / P.n_tiles;
                const int ms = r % P.m_super, b = r / P.m_super;
                const int row0 = ms * (TC2_MT * TC_BLOCK_M);
                const int n_seg = P.n_seg[n_tile];
                for (int s = 0; s < n_seg; ++s) {
                    const TcSeg sg = P.seg[n_tile][s];
                    if (sg.a_new) {
                        mbar_wait(a_empty(ia), pa ^ 1);
                        mbar_expect_tx(a_full(ia), TC2_MT * TC2_A_BYTES);
#pragma unroll
                        for (int mt = 0; mt < TC2_MT; ++mt)
                            tma_load_3d(sA + (ia * TC2_MT + mt) * TC2_A_SLOT, sg.src ? &tm_a1 : &tm_a0, a_full(ia), sg.col,
                                        row0 + mt * TC_BLOCK_M + sg.load_shift, b);
                        if (++ia == (uint32_t)SA) { ia = 0; pa ^= 1; }
                    }
                    mbar_wait(b_empty(ib), pb ^ 1);
                    mbar_expect_tx(b_full(ib), (uint32_t)sg.n_cnt * TC_BLOCK_K * 2);
                    const int wrow = n_tile * P.bn + sg.n_off;
                    for (int j = 0; j < sg.n_cnt; j += 64)
                        tma_load_2d(sB + ib * b_bytes + (uint32_t)j * 128, &tm_w, b_full(ib), sg.wk, wrow + j);
                    if (++ib == (uint32_t)SB) { ib = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t ia = 0, pa = 0, ib = 0, pb = 0, cur_a = 0;
            const uint64_t desc_hi = make_sw128_desc(0, 0);          // SBO, version, swizzle fields; address field empty
            int it = 0;
            for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
                const int n_tile = tile % P.n_tiles;
                const int n_seg = P.n_seg[n_tile];
                const int as = it % NACC;
                mbar_wait(acc_empty(as), (((uint32_t)(it / NACC)) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t acc0 = tmem_base + (uint32_t)(as * TC2_MT * P.bn);
                for (int s = 0; s < n_seg; ++s) {
                    const TcSeg sg = P.seg[n_tile][s];
                    if (sg.a_new) {
                        cur_a = ia;
                        mbar_wait(a_full(ia), pa);
                        if (++ia == (uint32_t)SA) { ia = 0; pa ^= 1; }
                    }
                    mbar_wait(b_full(ib), pb);
                    tc_fence_after();
                    // The single issuing thread is the pacing resource for N <= 128 (an MMA retires in 32-64 cycles), so the
                    // per-MMA instruction count is kept minimal: descriptors are a constant high word plus (address >> 4),
                    // and stepping K by 16 elements is "+2" in that field (smem addresses < 256 KB fit its 14 bits).
                    const uint32_t idesc = make_idesc(TC_BLOCK_M, (uint32_t)sg.n_cnt);
                    const uint64_t bd0 = desc_hi | (uint64_t)((sB + ib * b_bytes) >> 4);
                    const uint32_t acc_first = s > 0 ? 1u : 0u;
#pragma unroll
                    for (int mt = 0; mt < TC2_MT; ++mt) {
                        const uint32_t a0 = sA + (cur_a * TC2_MT + mt) * TC2_A_SLOT + (uint32_t)sg.desc_row * 128u;
                        uint64_t ad0 = desc_hi | (uint64_t)(a0 >> 4);
                        if (P.use_base_off) ad0 |= (uint64_t)((a0 >> 7) & 7u) << 49;
                        const uint32_t d = acc0 + (uint32_t)(mt * P.bn + sg.n_off);
                        umma_bf16(d, ad0, bd0, idesc, acc_first);
                        umma_bf16(d, ad0 + 2, bd0 + 2, idesc, 1u);
                        umma_bf16(d, ad0 + 4, bd0 + 4, idesc, 1u);
                        umma_bf16(d, ad0 + 6, bd0 + 6, idesc, 1u);
                    }
                    umma_commit(b_empty(ib));
                    if (sg.a_last) umma_commit(a_empty(cur_a));
                    if (++ib == (uint32_t)SB) { ib = 0; pb ^= 1; }
                }
                umma_commit(acc_full(as));
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int e = warp - 2;
        const int q = warp & 3;                       // TMEM lane quarter
        const int ch = e >> 2;                        // which half of the columns
        const int cols_per_warp = P.bn >> 1;
        const uint32_t stg0 = sStage + (uint32_t)e * NBUF * 4096u;
        const int cg = P.cout >> 3;
        const int n_part = P.m_tiles * P.n_tiles;
        int buf = 0, stores = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
            const int n_tile = tile % P.n_tiles;
            const int r = tile / P.n_tiles;
            const int ms = r % P.m_super, b = r / P.m_super;
            const int as = it % NACC;
            float* my_stat = s_stat + (((it & 1) * TC2_MT) * 8 + e) * 16;      // + mt * 8 * 16
            if (lane < 16) {
                my_stat[lane] = 0.0f;
                my_stat[8 * 16 + lane] = 0.0f;
            }
            __syncwarp();
            mbar_wait(acc_full(as), ((uint32_t)(it / NACC)) & 1u);
            tc_fence_after();
            for (int mt = 0; mt < TC2_MT; ++mt) {
                const int m_tile = ms * TC2_MT + mt;
                if (m_tile >= P.m_tiles) break;
                const int row_base = m_tile * TC_BLOCK_M + q * 32;
                const bool row_ok = row_base + lane < P.rows;
                float* st_mt = my_stat + mt * 8 * 16;
                const uint32_t acc = tmem_base + (uint32_t)((as * TC2_MT + mt) * P.bn) + ((uint32_t)(q * 32) << 16);
                for (int c0 = ch * cols_per_warp; c0 < (ch + 1) * cols_per_warp; c0 += 64) {
                    const uint32_t stg = stg0 + buf * 4096;
                    if (stores >= NBUF) {              // the staging buffer must have been read by its previous store
                        if (lane == 0) {
                            if (NBUF == 2) tma_wait_read<1>(); else tma_wait_read<0>();
                        }
                        __syncwarp();
                    }
                    {
                        const int ch0 = (n_tile * P.bn + c0) & (P.cout - 1);
                        epi_chunk64<CG_LOG2>(acc + (uint32_t)c0, s_bias + n_tile * P.bn + c0, row_ok, part != nullptr, st_mt,
                                             ch0 >> CG_LOG2, stg, lane);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(&tm_out, stg, n_tile * P.bn + c0, row_base, b);
                        tma_commit();
                    }
                    ++stores;
                    if (NBUF == 2) buf ^= 1;
                }
            }
            // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp, publish the stats
            tc_fence_before();
            named_bar_sync(1, 256);
            if (e == 0 && lane == 0) mbar_arrive(acc_empty(as));
            const int tid = threadIdx.x - 64;
            if (tid < TC2_MT * 8 && part != nullptr) {
                const int mt = tid >> 3, g = tid & 7;
                const int m_tile = ms * TC2_MT + mt;
                if (m_tile < P.m_tiles) {
                    const float* sp = s_stat + (((it & 1) * TC2_MT + mt) * 8) * 16 + g * 2;
                    float a1 = 0.0f, a2 = 0.0f;
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        a1 += sp[w * 16 + 0];
                        a2 += sp[w * 16 + 1];
                    }
                    float* pt = part + ((size_t)b * n_part + (m_tile * P.n_tiles + n_tile)) * 16;
                    pt[g * 2 + 0] = a1;
                    pt[g * 2 + 1] = a2;
                }
            }
        }
        if (lane == 0) tma_wait_all<0>();
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------ host side
static int build_params(const gw_conv_tc_shape* s, TcParams* P, bool halo = false) {
    memset(P, 0, sizeof(*P));
    GW_REQUIRE(s->Cout % 64 == 0 && s->Cout <= 256, "conv_tc: Cout=%d must be 64/128/192/256", s->Cout);
    GW_REQUIRE(s->C0 % 64 == 0 && s->C0 > 0 && s->C1 % 64 == 0, "conv_tc: C0=%d C1=%d must be multiples of 64", s->C0, s->C1);
    GW_REQUIRE((s->Cout & (s->Cout - 1)) == 0, "conv_tc: Cout=%d must be a power of two", s->Cout);
    P->cout = s->Cout;
    if (!s->pair) {
        GW_REQUIRE(s->n_src == 1 && s->C1 == 0 && s->L0 == s->L, "conv_tc: non-pair conv takes one source at the output length");
        P->rows = s->L;
        P->n_tiles = 1;
        P->bn = s->Cout;
        int n = 0;
        for (int c = 0; c < s->C0 / 64; ++c)
            for (int k = 0; k < 3; ++k) {
                GW_REQUIRE(n < TC_MAX_SEG, "conv_tc: too many segments");
                TcSeg& g = P->seg[0][n];
                g.src = 0; g.shift = (int16_t)(k - 1); g.col = (int16_t)(c * 64); g.n_off = 0; g.n_cnt = (int16_t)s->Cout;
                g.ci0 = (int16_t)(c * 64); g.mask[0] = (uint8_t)(1 << k); g.mask[1] = 0; g.wk = n * 64;
                ++n;
            }
        P->n_seg[0] = n;
    } else if (s->pair == 2) {
        // dgrad through the nearest upsample (pair-sum space): src0 = d_raw viewed as [B, L, 2*C0] with L0 = 2L rows,
        //   d_h[m] = P[m-1].hi A0 + P[m].lo (A0+A1) + P[m].hi (A1+A2) + P[m+1].lo A2,  A_k = taps of the dgrad weights
        GW_REQUIRE(s->n_src == 1 && s->C1 == 0 && s->L0 == 2 * s->L, "conv_tc: pair-sum dgrad takes one source with L0 = 2L");
        P->rows = s->L;
        P->n_tiles = 1;
        P->bn = s->Cout;
        int n = 0;
        for (int c = 0; c < s->C0 / 64; ++c) {
            const int col[4] = {c * 64, c * 64, s->C0 + c * 64, s->C0 + c * 64};
            const int shift[4] = {0, +1, -1, 0};
            const int mask[4] = {0b011, 0b100, 0b001, 0b110};
            for (int j = 0; j < 4; ++j) {
                GW_REQUIRE(n < TC_MAX_SEG, "conv_tc: too many segments");
                TcSeg& g = P->seg[0][n];
                g.src = 0; g.shift = (int16_t)shift[j]; g.col = (int16_t)col[j]; g.n_off = 0; g.n_cnt = (int16_t)s->Cout;
                g.ci0 = (int16_t)(c * 64); g.mask[0] = (uint8_t)mask[j]; g.mask[1] = 0; g.wk = n * 64;
                ++n;
            }
        }
        P->n_seg[0] = n;
    } else {
        // pair space: row m stands for output positions (2m, 2m+1), N = (phase, cout).
        //   pair == 1: decoder conv over cat[nearest-upsample(src0), src1];
        //   pair == 3: plain conv of ONE source evaluated in pair space (fills a 128-column tile when Cout = 64)
        if (s->pair == 1)
            GW_REQUIRE(s->n_src == 2 && s->C1 > 0 && s->L % 2 == 0 && s->L0 == s->L / 2,
                       "conv_tc: pair conv needs L even, L0 = L/2 and a skip source");
        else
            GW_REQUIRE(s->pair == 3 && s->n_src == 1 && s->C1 == 0 && s->L % 2 == 0 && s->L0 == s->L && 2 * s->Cout <= 256,
                       "conv_tc: pair-space plain conv needs one source, L even, Cout <= 128");
        P->rows = s->L / 2;
        const int ntot = 2 * s->Cout;
        P->n_tiles = ntot > 256 ? 2 : 1;
        P->bn = ntot / P->n_tiles;
        // candidate segments: (src, shift, col, ci0, mask phase0, mask phase1); the first one covers both phases
        struct Cand { int src, shift, col, ci0, m0, m1; };
        Cand cand[96];
        int nc = 0;
        if (s->pair == 1) {
            for (int c = 0; c < s->C0 / 64; ++c) {    // upsampled half: h[m-1], h[m], h[m+1]
                cand[nc++] = {0, 0, c * 64, c * 64, 0b110, 0b011};
                cand[nc++] = {0, -1, c * 64, c * 64, 0b001, 0};
                cand[nc++] = {0, +1, c * 64, c * 64, 0, 0b100};
            }
        }
        // pair view [lo = x[2m] | hi = x[2m+1]] of the skip (pair == 1, src1) or of the only source (pair == 3, src0)
        const int psrc = s->pair == 1 ? 1 : 0;
        const int pC = s->pair == 1 ? s->C1 : s->C0;
        const int pci0 = s->pair == 1 ? s->C0 : 0;
        for (int c = 0; c < pC / 64; ++c) {
            const int ci = pci0 + c * 64;
            cand[nc++] = {psrc, 0, c * 64, ci, 0b010, 0b001};           // S[m].lo : W1 | W0
            cand[nc++] = {psrc, +1, c * 64, ci, 0, 0b100};              // S[m+1].lo : -  | W2
            cand[nc++] = {psrc, 0, pC + c * 64, ci, 0b100, 0b010};      // S[m].hi : W2 | W1
            cand[nc++] = {psrc, -1, pC + c * 64, ci, 0b001, 0};         // S[m-1].hi : W0 | -
        }
        for (int t = 0; t < P->n_tiles; ++t) {
            int n = 0;
            for (int i = 0; i < nc; ++i) {
                const Cand& c = cand[i];
                int n_off, n_cnt;
                uint8_t m0 = (uint8_t)c.m0, m1 = (uint8_t)c.m1;
                if (P->n_tiles == 2) {                 // tile t = phase t
                    const int m = t == 0 ? c.m0 : c.m1;
                    if (!m) continue;
                    n_off = 0; n_cnt = s->Cout; m0 = (uint8_t)m; m1 = 0;
                } else if (c.m0 && c.m1) { n_off = 0; n_cnt = 2 * s->Cout; }
                else if (c.m0) { n_off = 0; n_cnt = s->Cout; }
                else { n_off = s->Cout; n_cnt = s->Cout; m0 = (uint8_t)c.m1; m1 = 0; }
                GW_REQUIRE(n < TC_MAX_SEG, "conv_tc: too many segments (%d)", n);
                TcSeg& g = P->seg[t][n];
                g.src = (int16_t)c.src; g.shift = (int16_t)c.shift; g.col = (int16_t)c.col;
                g.n_off = (int16_t)n_off; g.n_cnt = (int16_t)n_cnt; g.ci0 = (int16_t)c.ci0;
                g.mask[0] = m0; g.mask[1] = m1; g.wk = n * 64;
                ++n;
            }
            P->n_seg[t] = n;
        }
    }
    P->m_tiles = gw_cdiv(P->rows, TC_BLOCK_M);
    int mx = P->n_seg[0] > P->n_seg[1] ? P->n_seg[0] : P->n_seg[1];
    P->k_total = mx * 64;
    // v2: group consecutive segments on the same (src, col) around one halo A tile
    for (int t = 0; t < P->n_tiles; ++t) {
        for (int i = 0; i < P->n_seg[t]; ++i) {
            TcSeg& g = P->seg[t][i];
            const bool first = i == 0 || P->seg[t][i - 1].src != g.src || P->seg[t][i - 1].col != g.col;
            const bool last = i + 1 == P->n_seg[t] || P->seg[t][i + 1].src != g.src || P->seg[t][i + 1].col != g.col;
            if (halo) {
                g.a_new = first; g.a_last = last; g.load_shift = -1; g.desc_row = (int8_t)(g.shift + 1);
            } else {
                g.a_new = 1; g.a_last = 1; g.load_shift = (int8_t)g.shift; g.desc_row = 0;
            }
        }
    }
    P->m_super = gw_cdiv(P->m_tiles, 2);
    P->total_tiles = s->B * P->m_super * P->n_tiles;
    return GW_OK;
}

extern "C" long gw_conv_tc_packed_elems(const gw_conv_tc_shape* s) {
    TcParams P;
    if (build_params(s, &P) != GW_OK) return -1;
    return (long)P.n_tiles * P.bn * P.k_total;
}
extern "C" int gw_conv_tc_n_part(const gw_conv_tc_shape* s) {
    TcParams P;
    if (build_params(s, &P) != GW_OK) return -1;
    return P.m_tiles * P.n_tiles;
}

// packed[(t*bn + n)][seg*64 + kk] = sum over taps in mask of w[co][ci0 + kk][tap]
__global__ void conv_tc_pack_kernel(const __grid_constant__ TcParams P, const float* __restrict__ w, int cin,
                                    bf16* __restrict__ packed) {
    pdl_wait();
    pdl_launch_dependents();
    const int t = blockIdx.z, s = blockIdx.y;
    const TcSeg sg = P.seg[t][s];
    const bool active = s < P.n_seg[t];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.bn * 64; i += gridDim.x * blockDim.x) {
        const int n = i / 64, kk = i % 64;
        float v = 0.0f;
        if (active && n >= sg.n_off && n < sg.n_off + sg.n_cnt) {
            const int rel = n - sg.n_off;
            const int ph = rel / P.cout;             // 0 unless the segment spans both phases
            const int co = rel % P.cout;
            const int mask = sg.mask[ph];
            const float* wp = w + ((size_t)co * cin + sg.ci0 + kk) * 3;
            if (mask & 1) v += wp[0];
            if (mask & 2) v += wp[1];
            if (mask & 4) v += wp[2];
        }
        packed[((size_t)(t * P.bn + n)) * P.k_total + s * 64 + kk] = __float2bfloat16_rn(v);
    }
}

extern "C" int gw_conv_tc_pack(const gw_conv_tc_shape* s, const float* w, void* packed, void* stream) {
    TcParams P;
    int rc = build_params(s, &P);
    if (rc != GW_OK) return rc;
    dim3 grid(gw_cdiv(P.bn * 64, 256), P.k_total / 64, P.n_tiles);
    GW_CUDA(gw_launch_pdl(conv_tc_pack_kernel, grid, dim3(256), (size_t)(0), (cudaStream_t)stream, P, w, s->C0 + s->C1, (bf16*)packed));
    GW_LAUNCH_CHECK();
    return GW_OK;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// bf16 tensor [d2][d1][d0] (d0 contiguous), box [b0][b1][1], 128B swizzle, zero OOB fill
static int make_map3(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1) {
    PFN_encodeTiled enc = get_encode();
    GW_REQUIRE(enc != nullptr, "conv_tc: cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GW_REQUIRE(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(3d) failed with %d (dims %llu %llu %llu)", (int)r,
               (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2);
    return GW_OK;
}
static int make_map2(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint32_t b0, uint32_t b1) {
    PFN_encodeTiled enc = get_encode();
    GW_REQUIRE(enc != nullptr, "conv_tc: cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {d0 * 2};
    cuuint32_t box[2] = {b0, b1};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GW_REQUIRE(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(2d) failed with %d", (int)r);
    return GW_OK;
}

static int sm_count_cached() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    }
    return n;
}

// variant bits: 1 = v1 kernel sized for two CTAs/SM; 2 = v2 persistent super-tile kernel; 4 = (v2) halo-shared A tiles;
//               8 = (v2 halo) set the descriptor base_offset field from the start address
extern "C" int gw_conv_tc(const gw_conv_tc_shape* s, const void* src0, const void* src1, const void* packed,
                          const float* bias, void* raw, float* part, int variant, void* stream) {
    TcParams P;
    const bool v2 = (variant & 2) != 0;
    int rc = build_params(s, &P, v2 && (variant & 4));
    if (rc != GW_OK) return rc;
    GW_REQUIRE(src0 != nullptr && packed != nullptr && raw != nullptr, "conv_tc: null pointer");
    GW_REQUIRE((s->n_src == 2) == (src1 != nullptr), "conv_tc: src1 / n_src mismatch");
    GW_REQUIRE(!v2 || P.bn >= 128, "conv_tc v2 needs bn >= 128");
    const uint32_t a_box_rows = v2 ? 130 : TC_BLOCK_M;
    CUtensorMap ta0, ta1, tw, to;
    if (s->pair == 2 || s->pair == 3) {
        if ((rc = make_map3(&ta0, src0, (uint64_t)2 * s->C0, (uint64_t)P.rows, (uint64_t)s->B, 64, a_box_rows)) != GW_OK) return rc;
    } else {
        if ((rc = make_map3(&ta0, src0, (uint64_t)s->C0, (uint64_t)s->L0, (uint64_t)s->B, 64, a_box_rows)) != GW_OK) return rc;
    }
    if (s->pair == 1) {
        if ((rc = make_map3(&ta1, src1, (uint64_t)2 * s->C1, (uint64_t)P.rows, (uint64_t)s->B, 64, a_box_rows)) != GW_OK) return rc;
        if ((rc = make_map3(&to, raw, (uint64_t)2 * s->Cout, (uint64_t)P.rows, (uint64_t)s->B, 64, 32)) != GW_OK) return rc;
    } else {
        ta1 = ta0;
        const uint64_t oc = s->pair == 3 ? 2 * (uint64_t)s->Cout : (uint64_t)s->Cout;
        if ((rc = make_map3(&to, raw, oc, (uint64_t)P.rows, (uint64_t)s->B, 64, 32)) != GW_OK) return rc;
    }
    if ((rc = make_map2(&tw, packed, (uint64_t)P.k_total, (uint64_t)P.n_tiles * P.bn, 64, 64)) != GW_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int cg = s->Cout / 8;
    if (v2) {
        P.use_base_off = (variant & 8) ? 1 : 0;
        P.sa = 2;
        if (P.bn > 128) { P.sb = 3; P.nbuf = 1; P.n_acc = 1; }
        else { P.sb = 4; P.nbuf = 2; P.n_acc = 2; }
        const int smem = 1024 + P.sa * TC2_MT * TC2_A_SLOT + P.sb * P.bn * 128 + 8 * P.nbuf * 4096 + 256 + 2048 + 2048 + 64;
        GW_REQUIRE(smem <= 232448, "conv_tc v2: smem %d too large", smem);
        int grid = P.total_tiles < sm_count_cached() ? P.total_tiles : sm_count_cached();
#define TC2_GO(LG)                                                                                                  \
    do {                                                                                                            \
        GW_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));       \
        GW_CUDA(gw_launch_pdl(conv_tc2_kernel<LG>, grid, dim3(320), (size_t)(smem), st, ta0, ta1, tw, to, P, bias, part));                               \
    } while (0)
        if (cg == 8) TC2_GO(3);
        else if (cg == 16) TC2_GO(4);
        else TC2_GO(5);
#undef TC2_GO
        GW_LAUNCH_CHECK();
        return GW_OK;
    }
    const int stage_bytes = TC_BLOCK_M * TC_BLOCK_K * 2 + P.bn * TC_BLOCK_K * 2;
    const int fixed = 1024 /*align slack*/ + 4 * 2 * 4096 /*store staging*/ + 256 + 256 * 4 + 64 * 4 + 64;
    int max_seg = P.n_seg[0] > P.n_seg[1] ? P.n_seg[0] : P.n_seg[1];
    int budget = 232448;
    if (variant & 1) budget = 232448 / 2 - 1024;       // aim for two CTAs per SM
    int stages = (budget - fixed) / stage_bytes;
    if (stages > max_seg) stages = max_seg;
    if (stages > 8) stages = 8;
    GW_REQUIRE(stages >= 2 || max_seg == 1, "conv_tc: not enough shared memory for 2 stages (bn=%d)", P.bn);
    P.stages = stages;
    const int smem = fixed + stages * stage_bytes;
    dim3 grid(P.m_tiles * s->B, P.n_tiles);
#define TC_GO(LG)                                                                                                   \
    do {                                                                                                            \
        GW_CUDA(cudaFuncSetAttribute(conv_tc_kernel<LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));        \
        GW_CUDA(gw_launch_pdl(conv_tc_kernel<LG>, grid, dim3(192), (size_t)(smem), st, ta0, ta1, tw, to, P, bias, part));                                \
    } while (0)
    if (cg == 8) TC_GO(3);
    else if (cg == 16) TC_GO(4);
    else TC_GO(5);
#undef TC_GO
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ------------------------------------------------------------------------------------------------ fused conv + GroupNorm block
#include "conv_gn.cuh"
