// CFG combine + DDIM / DDPM update shared by every head kernel (stream_gn.cu, generic.cu): inference.py:455-506.
#pragma once
#include "common.cuh"

struct FssArgs {
    int mode, cfg_both, selfcond, pred_x0;
    float eps_scale, dc_weight;
    const float* y_dc;
    unsigned long long seed;
    long sample0;
    const unsigned long long* rng;      // device {seed, sample0} (graph-replay safe) or NULL -> the by-value fields
    unsigned int* advance;      // final_step_dots_kernel: CTA counter; the last CTA out does *step_ptr += 1 (NULL: nobody does)
};

struct FssCoef {
    float c_s1mab, c_sab, c_sabp, c_dir, c_sig, c_w, c_s1mab_cl;
    int use, last, draw;
};
// CFG combine + DDIM / DDPM update of one position (inference.py:455-506) given the head output of the conditional (outv0) and
// unconditional (outv1) rows; writes x_{t-1} / x0_hat into the other ping-pong buffer
__device__ __forceinline__ void fss_update(const FssArgs& p, const FssCoef& cf, float outv0, float outv1, float xt_c, float zu,
                                           const float* __restrict__ noise, float* __restrict__ net_out, int n_half, int b, int B,
                                           int Cx, int L, int l, float* __restrict__ eps_out, float* __restrict__ x0_out) {
    const float c_s1mab = cf.c_s1mab, c_sab = cf.c_sab, c_sabp = cf.c_sabp, c_dir = cf.c_dir, c_sig = cf.c_sig, c_w = cf.c_w;
    const float c_s1mab_cl = cf.c_s1mab_cl;
    const int use = cf.use, last = cf.last, draw = cf.draw;
    {
        float o;
        if (use == 0) o = outv0;
        else if (use == 1) o = p.cfg_both ? outv1 : outv0;
        else o = __fadd_rn(outv1, __fmul_rn(c_w, __fsub_rn(outv0, outv1)));
        float eps, x0;
        if (!p.pred_x0) {
            eps = __fmul_rn(p.eps_scale, o);
            x0 = __fdiv_rn(__fsub_rn(xt_c, __fmul_rn(c_s1mab, eps)), c_sab);
        } else {
            x0 = o;
            eps = __fdiv_rn(__fsub_rn(xt_c, __fmul_rn(c_sab, x0)), c_s1mab_cl);
        }
        if (p.dc_weight > 0.0f)
            x0 = __fadd_rn(__fmul_rn(1.0f - p.dc_weight, x0), __fmul_rn(p.dc_weight, p.y_dc[(size_t)b * L + l]));
        float xn;
        if (last) {
            xn = x0;
        } else {
            float nz = 0.0f;
            if (c_sig > 0.0f) {
                const float z = noise != nullptr ? noise[((size_t)draw * B + b) * L + l] : zu;
                nz = __fmul_rn(c_sig, z);
            }
            xn = __fadd_rn(__fadd_rn(__fmul_rn(c_sabp, x0), __fmul_rn(c_dir, eps)), nz);
        }
        for (int hf = 0; hf < n_half; ++hf) {
            float* orow = net_out + (size_t)(b + hf * B) * Cx * L;
            orow[l] = xn;
            if (p.selfcond) orow[(size_t)(Cx - 1) * L + l] = x0;
        }
        if (eps_out != nullptr) eps_out[(size_t)b * L + l] = eps;
        if (x0_out != nullptr) x0_out[(size_t)b * L + l] = x0;
    }
}

