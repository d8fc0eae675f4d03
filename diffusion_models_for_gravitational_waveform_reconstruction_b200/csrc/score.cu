// On-device scoring of reconstructions: the evaluation loops of the reference score every sample on
// the host in numpy (inference.py:11-27, 247-262, 303-314; sweep_infer.py:8-13, 225-241).  One CTA scores one sample, all
// accumulations in fp64 like the reference's np.float64 arithmetic.
//
// out[b, 0..11] (fp64):
//   0 corr_last   Pearson correlation over the tail window t >= t_max - secs        (_score_last_window / _corr)
//   1 mae_last    mean |x - c| over the same window
//   2 nmae_sigma  mean |x - c| over the last int(fs*secs) samples / (sigma + 1e-12)   (sweep_infer.py:232-237)
//   3 overlap     <x, c> / (|x| |c|) over the whole segment
//   4 best_lag    argmax_k sum_i c[i] x[i+k], |k| <= max_shift, first maximum wins    (_best_lag_by_xcorr(clean, xhat))
//   5 xc_mae      mean |x_al - c_al| for t in [-80 ms, +40 ms] around the clean peak after aligning by best_lag (_align_xcorr)
//   6 xc_nmae_clean = xc_mae / (mean |c_al| over the window + 1e-12)
//   7 xc_nmae_sigma = xc_mae / (sigma + 1e-12)
//   8 peak index of |c_al| (in aligned coordinates), 9 aligned length, 10 window count, 11 tail-window count
#include "common.cuh"
#include "../../include/gwb200.h"

#define SC_NOUT 12

__device__ __forceinline__ double block_sum_d(double v, double* red) {
    v = warp_sum_d(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
}

// Lag search out of shared memory, register-tiled: c and x staged ONCE as fp64 (x zero-padded by `pad` on both sides, so every
// (i, k) pair is in range and out-of-range products are exact zeros), a warp takes 8 consecutive lags, a lane 4 consecutive
// positions per iteration: 32 DFMA on 4 + 11 shared-memory loads instead of 2 loads (and 2 conversions) per DFMA.  Both series
// are stored in four planes (element e -> plane e & 3, slot e >> 2), so that lanes reading elements 4 apart touch consecutive
// 8-byte words: every load is bank-conflict free.
#define SC_LAGS 8
__device__ __forceinline__ void lag_search_tiled(const double* __restrict__ cs, const double* __restrict__ xs, int pad, int Lp, int nx,
                                                 int ms, int warp, int lane, double& bestv, int& bestk) {
    const int cq = Lp >> 2, xq = nx >> 2;
    for (int k0 = -ms + warp * SC_LAGS; k0 <= ms; k0 += 8 * SC_LAGS) {
        double acc[SC_LAGS];
#pragma unroll
        for (int j = 0; j < SC_LAGS; ++j) acc[j] = 0.0;
        const double* xp[SC_LAGS + 3];
#pragma unroll
        for (int j = 0; j < SC_LAGS + 3; ++j) {
            const int e = pad + k0 + j;                        // element of lane 0 at i = 0
            xp[j] = xs + (e & 3) * xq + (e >> 2) + lane;
        }
        for (int t = 0; t < cq; t += 32) {                    // positions i = 4 (lane + t) .. + 3
            double cv[4], xw[SC_LAGS + 3];
#pragma unroll
            for (int q = 0; q < 4; ++q) cv[q] = cs[q * cq + lane + t];
#pragma unroll
            for (int j = 0; j < SC_LAGS + 3; ++j) xw[j] = xp[j][t];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < SC_LAGS; ++j) acc[j] = fma(cv[q], xw[q + j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < SC_LAGS; ++j) {
            const double v = warp_sum_d(acc[j]);
            const int k = k0 + j;
            if (k <= ms && v > bestv) { bestv = v; bestk = k; }      // ascending k per warp: first maximum wins
        }
    }
}

__global__ void __launch_bounds__(256) score_kernel(const float* __restrict__ xhat, const float* __restrict__ clean,
                                                    const float* __restrict__ sigma, int L, double fs, double secs, int max_shift,
                                                    double delta_t, double* __restrict__ out, int tiled_pad) {
    extern __shared__ __align__(16) double sc_dyn[];          // tiled lag search: cs [Lp] | xs [pad + Lp + pad + 16]
    __shared__ double red[8];
    __shared__ double s_best[8];
    __shared__ int s_bestk[8];
    __shared__ int s_pk;
    const int b = blockIdx.x;
    const float* x = xhat + (size_t)b * L;
    const float* c = clean + (size_t)b * L;
    const double sg = sigma ? (double)sigma[b] : 1.0;
    double* o = out + (size_t)b * SC_NOUT;
    // ---- tail window: indices i with i/fs >= (L-1)/fs - secs   (inference.py:11-13)
    const double tmax = (double)(L - 1) / fs;
    int i0 = 0;
    {
        // smallest i with i/fs >= tmax - secs, evaluated exactly as numpy does (float64 division, comparison)
        int lo = 0, hi = L - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((double)mid / fs >= tmax - secs) hi = mid; else lo = mid + 1;
        }
        i0 = lo;
    }
    const int nwin = L - i0;
    double sx = 0.0, sc = 0.0, sae = 0.0;
    for (int i = i0 + threadIdx.x; i < L; i += 256) {
        sx += (double)x[i];
        sc += (double)c[i];
        sae += fabs((double)x[i] - (double)c[i]);
    }
    sx = block_sum_d(sx, red);
    sc = block_sum_d(sc, red);
    sae = block_sum_d(sae, red);
    const double mx = sx / nwin, mc = sc / nwin;
    double sxx = 0.0, scc = 0.0, sxc = 0.0;
    for (int i = i0 + threadIdx.x; i < L; i += 256) {
        const double a = (double)x[i] - mx, d = (double)c[i] - mc;
        sxx += a * a;
        scc += d * d;
        sxc += a * d;
    }
    sxx = block_sum_d(sxx, red);
    scc = block_sum_d(scc, red);
    sxc = block_sum_d(sxc, red);
    // ---- nmae over the last int(fs*secs) samples
    int w = (int)(fs * secs);
    if (w > L) w = L;
    double sae2 = 0.0;
    for (int i = L - w + threadIdx.x; i < L; i += 256) sae2 += fabs((double)x[i] - (double)c[i]);
    sae2 = block_sum_d(sae2, red);
    // ---- overlap over the whole segment
    double pxx = 0.0, pcc = 0.0, pxc = 0.0;
    for (int i = threadIdx.x; i < L; i += 256) {
        const double a = (double)x[i], d = (double)c[i];
        pxx += a * a;
        pcc += d * d;
        pxc += a * d;
    }
    pxx = block_sum_d(pxx, red);
    pcc = block_sum_d(pcc, red);
    pxc = block_sum_d(pxc, red);
    // ---- best lag: v(k) = sum_i c[i] * x[i + k]   (a = clean, b = xhat in _best_lag_by_xcorr(a, b))
    int ms = max_shift <= 0 ? L - 1 : max_shift;
    if (ms > L - 1) ms = L - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double bestv = -1.0e300;
    int bestk = 0;
    if (tiled_pad > 0) {
        const int Lp = (L + 127) / 128 * 128;
        double* cs = sc_dyn;
        double* xs = sc_dyn + Lp;
        const int nx = tiled_pad + Lp + tiled_pad + 16;
        for (int i = threadIdx.x; i < Lp; i += 256) cs[(i & 3) * (Lp >> 2) + (i >> 2)] = i < L ? (double)c[i] : 0.0;
        for (int i = threadIdx.x; i < nx; i += 256) {
            const int j = i - tiled_pad;
            xs[(i & 3) * (nx >> 2) + (i >> 2)] = (j >= 0 && j < L) ? (double)x[j] : 0.0;
        }
        __syncthreads();
        lag_search_tiled(cs, xs, tiled_pad, Lp, nx, ms, warp, lane, bestv, bestk);
    } else
    for (int k = -ms + warp; k <= ms; k += 8) {
        const int lo = k < 0 ? -k : 0, hi = k > 0 ? L - k : L;       // i range with 0 <= i + k < L
        double v = 0.0;
        for (int i = lo + lane; i < hi; i += 32) v = fma((double)c[i], (double)x[i + k], v);
        v = warp_sum_d(v);
        if (v > bestv) { bestv = v; bestk = k; }                       // ascending k per warp: first maximum wins
    }
    if (lane == 0) { s_best[warp] = bestv; s_bestk[warp] = bestk; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bv = s_best[0];
        int bk = s_bestk[0];
        for (int wv = 1; wv < 8; ++wv)
            if (s_best[wv] > bv || (s_best[wv] == bv && s_bestk[wv] < bk)) { bv = s_best[wv]; bk = s_bestk[wv]; }
        s_bestk[0] = bk;
    }
    __syncthreads();
    const int k = s_bestk[0];
    // ---- alignment (inference.py:264-279): a_al = c[start:stop], b_al = x[start+k:stop+k]
    int start = k < 0 ? -k : 0;
    int stop = L < L - k ? L : L - k;
    int kk = k;
    if (stop <= start) { start = 0; stop = L; kk = 0; }
    const int La = stop - start;
    // peak of |c_al| (first maximum)
    double pv = -1.0;
    int pi = 0;
    for (int i = threadIdx.x; i < La; i += 256) {
        const double a = fabs((double)c[start + i]);
        if (a > pv) { pv = a; pi = i; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, pv, off);
        const int oi = __shfl_xor_sync(0xffffffffu, pi, off);
        if (ov > pv || (ov == pv && oi < pi)) { pv = ov; pi = oi; }
    }
    if (lane == 0) { s_best[warp] = pv; s_bestk[warp] = pi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bv = s_best[0];
        int bi = s_bestk[0];
        for (int wv = 1; wv < 8; ++wv)
            if (s_best[wv] > bv || (s_best[wv] == bv && s_bestk[wv] < bi)) { bv = s_best[wv]; bi = s_bestk[wv]; }
        s_pk = bi;
    }
    __syncthreads();
    const int pk = s_pk;
    // window t in [-0.080, 0.040] with t_i = i*delta_t - pk*delta_t (float64, as numpy: t -= t[pk])
    double wae = 0.0, wac = 0.0, wn = 0.0;
    const double tpk = (double)pk * delta_t;
    for (int i = threadIdx.x; i < La; i += 256) {
        const double t = (double)i * delta_t - tpk;
        if (t >= -0.080 && t <= 0.040) {
            const double a = (double)c[start + i], r = (double)x[start + kk + i];
            wae += fabs(r - a);
            wac += fabs(a);
            wn += 1.0;
        }
    }
    wae = block_sum_d(wae, red);
    wac = block_sum_d(wac, red);
    wn = block_sum_d(wn, red);
    if (threadIdx.x == 0) {
        o[0] = sxc / (sqrt(sxx * scc) + 1e-30);
        o[1] = sae / nwin;
        o[2] = (w > 0 ? sae2 / w : 0.0) / (sg + 1e-12);
        o[3] = pxc / (sqrt(pxx) * sqrt(pcc) + 1e-30);
        o[4] = (double)k;
        const double mae = wn > 0.0 ? wae / wn : 0.0;
        o[5] = mae;
        o[6] = mae / ((wn > 0.0 ? wac / wn : 0.0) + 1e-12);
        o[7] = mae / (sg + 1e-12);
        o[8] = (double)pk;
        o[9] = (double)La;
        o[10] = wn;
        o[11] = (double)nwin;
    }
}

extern "C" int gw_score_batch(const float* xhat, const float* clean, const float* sigma, int B, int L, double fs, double secs,
                              int max_shift, double delta_t, double* out, void* stream) {
    GW_REQUIRE(B > 0 && L > 1 && fs > 0.0 && secs > 0.0 && xhat != nullptr && clean != nullptr && out != nullptr,
               "gw_score_batch: arguments");
    // shared-memory lag search when the fp64 copies of both series fit (L = 4096, |k| <= 82: 68 KB); else straight from L1 / L2
    int ms = max_shift <= 0 ? L - 1 : max_shift;
    if (ms > L - 1) ms = L - 1;
    const int Lp = (L + 127) / 128 * 128;
    int pad = (ms + SC_LAGS * 8 + 15) / 16 * 16;              // k0 + j may run SC_LAGS * 8 - 1 past ms in the last pass of a warp
    size_t smem = (size_t)(Lp + pad + Lp + pad + 16) * sizeof(double);
    if (smem > 200 * 1024) { pad = 0; smem = 0; }
    if (smem > 48 * 1024) GW_CUDA(cudaFuncSetAttribute(score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    score_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(xhat, clean, sigma, L, fs, secs, max_shift, delta_t, out, pad);
    GW_LAUNCH_CHECK();
    return GW_OK;
}
