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

// HBM-streaming bf16 implementations (stream_gn.cu); same partial layouts as the register-streaming kernels
int gn_bwd_stream_rows(int L, int C);
int gn_bwd_stats_stream(const GnBwdArgs& a, int B, float* partial, cudaStream_t st);
int gn_bwd_apply_stream(const GnBwdArgs& a, int B, const float* gstat, void* d_raw, float* partial_bias, cudaStream_t st);
