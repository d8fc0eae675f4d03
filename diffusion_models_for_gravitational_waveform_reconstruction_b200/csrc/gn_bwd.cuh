// Shared between backward.cu (dispatch, finalize kernels) and stream_gn.cu (HBM-streaming bf16 kernels).
#pragma once
#include "common.cuh"

struct GnBwdArgs {
    const void* raw;        // [B, L, C] conv output saved by the forward
    const float* stats;     // [B, 8, 2] (mean, rstd) saved by gw_gn_apply
    const float* gn_w;      // [C]
    const float* gn_b;      // [C]
    const float* cond;      // [B, L, Cc] fp32 or NULL
    const float* wc;        // [C, Cc]
    const float* bc;        // [C]
    const float* film;      // row of sample b: film + b*film_b_stride + film_off; gamma at [0,C), beta at [C,2C)
    long film_b_stride;
    int film_off;
    const void* do_a;       // [B, L, C] gradient wrt out, or NULL
    const void* do_pool;    // [B, L/2, C] gradient wrt the pooled output (encoders), or NULL
    const float* do_eps;    // [B, L] fp32: the block feeds the head conv; dout[l,c] = sum_k do_w[c,k] * do_eps[l-k+1] (or NULL)
    const float* do_w;      // [C+1, 3] head weights (models.py:230)
    int L, C, Cc;
    int rows_per_cta;
};

#ifdef __CUDACC__
__device__ __forceinline__ f32x2 bf2_lo(uint32_t w) { return pk2(w << 16, w & 0xffff0000u); }

// z/2, sigmoid and silu derivative of a channel pair from one tanh.approx per element
__device__ __forceinline__ void sg_silu_pair(f32x2 x, f32x2 hA, f32x2 hB, f32x2& z, f32x2& act, f32x2& dact) {
    const f32x2 one = pkf2(1.0f, 1.0f), half2 = pkf2(0.5f, 0.5f), neg1 = pkf2(-1.0f, -1.0f);
    const f32x2 hh = ffma2(x, hA, hB);
    z = fadd2(hh, hh);
    float h0, h1, t0, t1;
    upk2(hh, h0, h1);
    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
    const f32x2 sg = ffma2(pkf2(t0, t1), half2, half2);
    act = fmul2(z, sg);
    dact = fmul2(sg, ffma2(z, ffma2(sg, neg1, one), one));
}
#endif

// HBM-streaming bf16 implementations (stream_gn.cu); same partial layouts as the register-streaming kernels
int gn_bwd_stream_rows(int L, int C);
bool gn_bwd_stream_fast_ok(const GnBwdArgs& a);   // the compile-time-specialised kernels run
int gn_bwd_stats_stream(const GnBwdArgs& a, int B, float* partial, cudaStream_t st);
int gn_bwd_apply_stream(const GnBwdArgs& a, int B, const float* gstat, void* d_raw, float* partial_bias, cudaStream_t st);

// one-pass implementation (gn_bwd_fused.cu): a group of CTAs keeps a sample's operands in shared memory
int gn_bwd_fused_group(int L, int C, int Cc, bool has_do, bool has_pool);
int gn_bwd_fused(const GnBwdArgs& a, int B, float* partial, float* partial_bias, void* d_raw, void* sync, cudaStream_t st);
