// GPU whitening / de-whitening and sigma estimation around the reverse chain.
// Reference (numpy, one sample at a time, float64): inference.py:125-205 -- _pick_sigma / _mad_std, _whiten_pair_train_like,
// _dewhiten_train_like, _whiten_pair_model, _dewhiten_model, _interp_psd_for_length; dataloader.py:110-151 uses the same
// train-like recipe.  Here: batched fp64 cuFFT (D2Z / Z2D) with hand-written spectral kernels between the transforms.
// Built as a separate library (libgwb200_fft.so) so that the core library carries no cuFFT dependency.
#include "common.cuh"
#include "../../include/gwb200_fft.h"
#include <cufft.h>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

static thread_local char g_ferr[512] = "ok";
void gw_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_ferr, sizeof(g_ferr), fmt, ap);
    va_end(ap);
}
extern "C" const char* gwf_last_error(void) { return g_ferr; }

#define GW_CUFFT(expr)                                                                      \
    do {                                                                                    \
        cufftResult _r = (expr);                                                            \
        if (_r != CUFFT_SUCCESS) {                                                          \
            gw_set_error("%s:%d %s -> cufft error %d", __FILE__, __LINE__, #expr, (int)_r); \
            return GW_ERR_CUDA;                                                             \
        }                                                                                   \
    } while (0)

// ---------------------------------------------------------------------------------------------- plans (cached per shape)
static std::mutex g_plan_mu;
static std::map<std::tuple<int, int, int>, cufftHandle> g_plans;      // (kind 0 = D2Z / 1 = Z2D, L, B)
static int get_plan(int kind, int L, int B, cufftHandle* out) {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto key = std::make_tuple(kind, L, B);
    auto it = g_plans.find(key);
    if (it == g_plans.end()) {
        cufftHandle h;
        int n[1] = {L};
        GW_CUFFT(cufftPlanMany(&h, 1, n, nullptr, 1, L, nullptr, 1, L / 2 + 1, kind == 0 ? CUFFT_D2Z : CUFFT_Z2D, B));
        it = g_plans.emplace(key, h).first;
    }
    *out = it->second;
    return GW_OK;
}

__device__ __forceinline__ double blk_sum_d(double v, double* red) {
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

// y (fp32) -> y64 - mean(y64) (or plain cast when demean == 0); one CTA per sample
__global__ void __launch_bounds__(256) to_f64_kernel(const float* __restrict__ y, int L, int demean, double* __restrict__ out) {
    __shared__ double red[8];
    const int b = blockIdx.x;
    double s = 0.0;
    if (demean) {
        for (int i = threadIdx.x; i < L; i += 256) s += (double)y[(size_t)b * L + i];
        s = blk_sum_d(s, red) / (double)L;
    }
    for (int i = threadIdx.x; i < L; i += 256) out[(size_t)b * L + i] = (double)y[(size_t)b * L + i] - s;
}

// P = max(conv_same(|Y|^2, ones(9)/9), 1e-20)   (inference.py:141-146: np.convolve(P, kernel, mode="same") when P.size > 9)
__global__ void __launch_bounds__(256) periodogram_kernel(const cufftDoubleComplex* __restrict__ Y, int F, double* __restrict__ P) {
    const int b = blockIdx.y;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const cufftDoubleComplex* yb = Y + (size_t)b * F;
    double acc = 0.0;
    if (F > 9) {
        // numpy accumulates the 9 products in index order of the full convolution: sum_j P[f + 4 - j] * k[j]
        for (int j = 0; j < 9; ++j) {
            const int i = f + 4 - j;
            if (i >= 0 && i < F) acc += (yb[i].x * yb[i].x + yb[i].y * yb[i].y) * (1.0 / 9.0);
        }
    } else {
        acc = yb[f].x * yb[f].x + yb[f].y * yb[f].y;
    }
    P[(size_t)b * F + f] = fmax(acc, 1e-20);
}

// Z = Y * g(P):  mode 0: 1/sqrt(P)   mode 1: 1/sqrt(P + 1e-12)   mode 2: sqrt(P + 1e-12)   mode 3: 1/sqrt(P + 1e-20)
__global__ void __launch_bounds__(256) spectral_scale_kernel(const cufftDoubleComplex* __restrict__ Y, const double* __restrict__ P,
                                                             long p_b_stride, int F, int mode, cufftDoubleComplex* __restrict__ Z) {
    const int b = blockIdx.y;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const double p = P[(size_t)b * p_b_stride + f];
    const cufftDoubleComplex y = Y[(size_t)b * F + f];
    cufftDoubleComplex z;
    if (mode == 2) {
        const double g = sqrt(p + 1e-12);
        z.x = y.x * g; z.y = y.y * g;
    } else {
        const double g = sqrt(mode == 0 ? p : (mode == 3 ? p + 1e-20 : p + 1e-12));
        z.x = y.x / g; z.y = y.y / g;
    }
    Z[(size_t)b * F + f] = z;
}

// irfft normalisation (1/L) + optional cast to fp32
__global__ void __launch_bounds__(256) finish_kernel(const double* __restrict__ t, long n, double scale, float* __restrict__ o32,
                                                     double* __restrict__ o64) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double v = t[i] * scale;
        if (o32) o32[i] = (float)v;
        if (o64) o64[i] = v;
    }
}

extern "C" long gwf_workspace_bytes(int B, int L) {
    const long F = L / 2 + 1;
    return (long)B * L * 8 + 2 * (long)B * F * 16;            // time-domain fp64 buffer + two spectra
}

static int fft_forward(const float* y, int demean, int B, int L, double* tbuf, cufftDoubleComplex* Y, cudaStream_t st) {
    to_f64_kernel<<<B, 256, 0, st>>>(y, L, demean, tbuf);
    GW_LAUNCH_CHECK();
    cufftHandle h;
    int rc = get_plan(0, L, B, &h);
    if (rc != GW_OK) return rc;
    GW_CUFFT(cufftSetStream(h, st));
    GW_CUFFT(cufftExecD2Z(h, tbuf, Y));
    return GW_OK;
}
static int fft_inverse(cufftDoubleComplex* Z, int B, int L, double* tbuf, float* o32, double* o64, cudaStream_t st) {
    cufftHandle h;
    int rc = get_plan(1, L, B, &h);
    if (rc != GW_OK) return rc;
    GW_CUFFT(cufftSetStream(h, st));
    GW_CUFFT(cufftExecZ2D(h, Z, tbuf));
    const long n = (long)B * L;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    finish_kernel<<<grid, 256, 0, st>>>(tbuf, n, 1.0 / (double)L, o32, o64);
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ---------------------------------------------------------------------------------------------- fused train-like whitening
// One CTA per sample, everything between the fp32 input and the fp32 output in shared memory (power-of-two L up to 8192):
// mean removal, the real FFT as an L/2-point complex fp64 FFT of the packed samples (radix-2 decimation in frequency in place,
// bit-reversal permutation, even / odd split), the 9-tap smoothed periodogram, division by sqrt(P), the inverse transform the same
// way back, for y and then for x with the same P.  HBM traffic = the algorithmic bytes (inputs read once, outputs written once);
// the cuFFT path above moves ~8x that through fp64 workspaces.
static std::map<int, double2*> g_twiddles;                  // L -> W_L^k = exp(-2 pi i k / L), k = 0 .. L/2 (device)
__global__ void twiddle_kernel(int L, double2* __restrict__ tab) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > L / 2) return;
    double s, c;
    sincospi(2.0 * (double)k / (double)L, &s, &c);
    tab[k] = make_double2(c, -s);
}
static int get_twiddles(int L, cudaStream_t st, const double2** out) {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto it = g_twiddles.find(L);
    if (it == g_twiddles.end()) {
        double2* tab = nullptr;
        GW_CUDA(cudaMalloc(&tab, (size_t)(L / 2 + 1) * sizeof(double2)));
        twiddle_kernel<<<gw_cdiv(L / 2 + 1, 256), 256, 0, st>>>(L, tab);
        GW_LAUNCH_CHECK();
        GW_CUDA(cudaStreamSynchronize(st));                  // one-time: the cached table may be used from any stream afterwards
        it = g_twiddles.emplace(L, tab).first;
    }
    *out = it->second;
    return GW_OK;
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

// Shared-memory layouts.  z (N complex fp64) and the twiddle table W_N^e (e < N/2) are stored skewed (one extra element per 8, per 64
// and per 512): the power-of-two strides of the late passes, of the twiddle exponents and of the bit-reversal permutation then still
// spread over the 8 sixteen-byte bank groups instead of piling onto one (measured: the unskewed permutation alone cost 4x the FFT).
__device__ __forceinline__ int zp(int i) { return i + (i >> 3) + (i >> 6) + (i >> 9); }
__device__ __forceinline__ int tw_skew(int i) { return i + (i >> 3) + (i >> 6) + (i >> 9); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }             // a * (-i)
template <int N>
__device__ __forceinline__ double2 tw_get(const double2* __restrict__ tws, int e) {                  // W_N^e, 0 <= e < N
    const double2 w = tws[tw_skew(e & (N / 2 - 1))];
    return (e & (N / 2)) ? make_double2(-w.x, -w.y) : w;
}
// forward 4-point DFT, outputs in natural order
__device__ __forceinline__ void dft4(double2 c0, double2 c1, double2 c2, double2 c3, double2& y0, double2& y1, double2& y2, double2& y3) {
    const double2 d0 = cadd(c0, c2), d1 = csub(c0, c2), d2 = cadd(c1, c3), d3 = mul_mi(csub(c1, c3));
    y0 = cadd(d0, d2); y2 = csub(d0, d2); y1 = cadd(d1, d3); y3 = csub(d1, d3);
}

// in-place forward N-point complex FFT of z (shared memory, zp-padded), N = 2^LGN, natural order in and out: decimation in
// frequency in radix-8 passes (a thread holds the 8 points of a butterfly in registers: N = 2048 is 8 x 8 x 8 x 4, four
// shared-memory round trips instead of eleven), one radix-4 or radix-2 pass for the leftover bits, then the bit-reversal
// permutation (every butterfly stores output q in sub-block bitrev(q), so the radix-2 ordering is kept).  The inverse transform is
// conj(forward(conj(.))), applied by the callers.  All 256 threads must call.
template <int LGN>
__device__ __forceinline__ void smem_fft(double2* z, const double2* __restrict__ tws) {
    constexpr int N = 1 << LGN;
    constexpr double RH = 0.70710678118654752440;
    int lgn = LGN;                                           // log2 of the current block size
#pragma unroll 1
    for (int pass = 0; pass < LGN / 3; ++pass) {
        const int m = 1 << (lgn - 3), sh = LGN - lgn;
#pragma unroll 1
        for (int t = threadIdx.x; t < N / 8; t += 256) {
            const int j = t & (m - 1), base = ((t - j) << 3) + j;
            double2 x[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) x[r] = z[zp(base + r * m)];
            // first layer: pairs (r, r + 4); the differences take W_8^r
            const double2 a0 = cadd(x[0], x[4]), a1 = cadd(x[1], x[5]), a2 = cadd(x[2], x[6]), a3 = cadd(x[3], x[7]);
            const double2 b0 = csub(x[0], x[4]);
            double2 b1 = csub(x[1], x[5]), b2 = csub(x[2], x[6]), b3 = csub(x[3], x[7]);
            b1 = make_double2(RH * (b1.x + b1.y), RH * (b1.y - b1.x));           // * (1 - i) / sqrt 2
            b2 = mul_mi(b2);
            b3 = make_double2(RH * (b3.y - b3.x), -RH * (b3.x + b3.y));          // * (-1 - i) / sqrt 2
            double2 y[8];
            dft4(a0, a1, a2, a3, y[0], y[2], y[4], y[6]);
            dft4(b0, b1, b2, b3, y[1], y[3], y[5], y[7]);
            // output q -> sub-block bitrev3(q), times W_n^(j q) = W_N^((j q) << sh): one table load (j << sh < N / 8), the other six
            // powers by multiplication (depth 3; the fp64 pipe has room, the shared-memory pipe is what this kernel runs into)
            double2 w[8];
            w[1] = tws[tw_skew(j << sh)];
            w[2] = cmul(w[1], w[1]);
            w[3] = cmul(w[2], w[1]);
            w[4] = cmul(w[2], w[2]);
            w[5] = cmul(w[4], w[1]);
            w[6] = cmul(w[4], w[2]);
            w[7] = cmul(w[4], w[3]);
            z[zp(base)] = y[0];
#pragma unroll
            for (int q = 1; q < 8; ++q) {
                const int slot = ((q & 1) << 2) | (q & 2) | (q >> 2);
                z[zp(base + slot * m)] = cmul(y[q], w[q]);
            }
        }
        __syncthreads();
        lgn -= 3;
    }
    if (LGN % 3 == 2) {                                      // blocks of 4 are left: j = 0, no twiddles
#pragma unroll 1
        for (int t = threadIdx.x; t < N / 4; t += 256) {
            const int base = t << 2;
            double2 y0, y1, y2, y3;
            dft4(z[zp(base)], z[zp(base + 1)], z[zp(base + 2)], z[zp(base + 3)], y0, y1, y2, y3);
            z[zp(base)] = y0; z[zp(base + 2)] = y1; z[zp(base + 1)] = y2; z[zp(base + 3)] = y3;
        }
        __syncthreads();
    } else if (LGN % 3 == 1) {
#pragma unroll 1
        for (int t = threadIdx.x; t < N / 2; t += 256) {
            const double2 a = z[zp(2 * t)], b = z[zp(2 * t + 1)];
            z[zp(2 * t)] = cadd(a, b);
            z[zp(2 * t + 1)] = csub(a, b);
        }
        __syncthreads();
    }
#pragma unroll 1
    for (int i = threadIdx.x; i < N; i += 256) {
        const int r = (int)(__brev((unsigned)i) >> (32 - LGN));
        if (i < r) {
            const double2 a = z[zp(i)];
            z[zp(i)] = z[zp(r)];
            z[zp(r)] = a;
        }
    }
    __syncthreads();
}

// z [N] (the packed-sample FFT) -> the real-input spectrum Y[0 .. N]: Y[k] (k < N) in place, Y[N] (real) to *yN
__device__ __forceinline__ void rfft_split(double2* z, int N, const double2* __restrict__ tw, double* yN) {
    for (int k = threadIdx.x; k <= N / 2; k += 256) {
        if (k == 0) {
            const double2 a = z[0];
            *yN = a.x - a.y;
            z[0] = make_double2(a.x + a.y, 0.0);
        } else if (2 * k == N) {
            z[zp(k)] = cconj(z[zp(k)]);
        } else {
            const double2 a = z[zp(k)], bc = cconj(z[zp(N - k)]);
            const double2 e = make_double2(0.5 * (a.x + bc.x), 0.5 * (a.y + bc.y));
            const double2 d = make_double2(0.5 * (a.x - bc.x), 0.5 * (a.y - bc.y));
            const double2 t = cmul(__ldg(tw + k), mul_mi(d));                    // W_L^k * (-i d)
            z[zp(k)] = make_double2(e.x + t.x, e.y + t.y);
            z[zp(N - k)] = make_double2(e.x - t.x, -(e.y - t.y));
        }
    }
    __syncthreads();
}
// inverse of rfft_split after the spectrum was multiplied by the real gains g[0 .. N]: Y -> the CONJUGATE of the packed-sample
// spectrum, in place (the forward FFT of it, conjugated again by the caller, is the inverse transform)
__device__ __forceinline__ void irfft_merge_conj(double2* z, int N, const double2* __restrict__ tw, const double* __restrict__ g,
                                                 double yN) {
    for (int k = threadIdx.x; k <= N / 2; k += 256) {
        if (k == 0) {
            const double y0 = z[0].x * g[0], yn = yN * g[N];                     // imaginary parts of Y[0], Y[N] are ignored (irfft)
            z[0] = make_double2(0.5 * (y0 + yn), -0.5 * (y0 - yn));
        } else if (2 * k == N) {
            const double2 a = z[zp(k)];
            z[zp(k)] = make_double2(a.x * g[k], a.y * g[k]);                     // conj(conj(Y[N/2]))
        } else {
            const double2 zk = z[zp(k)], zn = z[zp(N - k)];
            const double2 a = make_double2(zk.x * g[k], zk.y * g[k]);
            const double2 bc = make_double2(zn.x * g[N - k], -zn.y * g[N - k]);
            const double2 e = make_double2(0.5 * (a.x + bc.x), 0.5 * (a.y + bc.y));
            const double2 t = make_double2(0.5 * (a.x - bc.x), 0.5 * (a.y - bc.y));
            const double2 o = cmul(cconj(__ldg(tw + k)), t);
            z[zp(k)] = make_double2(e.x - o.y, -(e.y + o.x));                    // conj(E + i O)
            z[zp(N - k)] = make_double2(e.x + o.y, -(-e.y + o.x));               // conj(conj(E) + i conj(O))
        }
    }
    __syncthreads();
}

template <int LGN>
__global__ void __launch_bounds__(256, 2) whiten_fused_kernel(const float* __restrict__ y, const float* __restrict__ x,
                                                              const double2* __restrict__ tw, float* __restrict__ y_w,
                                                              float* __restrict__ x_w, double* __restrict__ P) {
    extern __shared__ __align__(16) unsigned char wf_smem[];
    __shared__ double red[8];
    __shared__ double s_yN;
    constexpr int N = 1 << LGN, L = 2 * N, F = N + 1;
    constexpr int MPT = (F + 255) / 256;                      // spectrum bins per thread
    const int b = blockIdx.x;
    double2* z = reinterpret_cast<double2*>(wf_smem);        // [N] skewed (x 1.143)
    double2* tws = z + (N + N / 6 + 8);                       // [N/2] skewed (x 1.143): W_N^e = tw[2e]
    double* g = reinterpret_cast<double*>(tws + (N / 2 + N / 12 + 8));   // [F]: |Y|^2, then 1 / sqrt(P)
    for (int j = threadIdx.x; j < N / 2; j += 256) tws[tw_skew(j)] = tw[2 * j];
#pragma unroll 1
    for (int pass = 0; pass < (x != nullptr ? 2 : 1); ++pass) {
        const float* src = (pass == 0 ? y : x) + (size_t)b * L;
        float* dst = (pass == 0 ? y_w : x_w) + (size_t)b * L;
        double s = 0.0;
        for (int i = threadIdx.x; i < L; i += 256) s += (double)src[i];
        s = blk_sum_d(s, red) / (double)L;
        for (int n = threadIdx.x; n < N; n += 256) {
            const float2 v = *reinterpret_cast<const float2*>(src + 2 * n);
            z[zp(n)] = make_double2((double)v.x - s, (double)v.y - s);
        }
        __syncthreads();
        smem_fft<LGN>(z, tws);
        rfft_split(z, N, tw, &s_yN);
        if (pass == 0) {
            for (int k = threadIdx.x; k < F; k += 256) {
                const double2 v = z[zp(k < N ? k : 0)];
                g[k] = k < N ? v.x * v.x + v.y * v.y : s_yN * s_yN;
            }
            __syncthreads();
            // P = max(conv_same(|Y|^2, ones(9) / 9), 1e-20), the products summed in numpy's order (see periodogram_kernel)
            double pv[MPT];
#pragma unroll
            for (int m = 0; m < MPT; ++m) {
                const int k = (int)threadIdx.x + 256 * m;
                double acc = 0.0;
                if (k < F) {
                    if (F > 9) {
#pragma unroll
                        for (int j = 0; j < 9; ++j) {
                            const int i = k + 4 - j;
                            if (i >= 0 && i < F) acc += g[i] * (1.0 / 9.0);
                        }
                    } else acc = g[k];
                }
                pv[m] = fmax(acc, 1e-20);
            }
            __syncthreads();
#pragma unroll
            for (int m = 0; m < MPT; ++m) {
                const int k = (int)threadIdx.x + 256 * m;
                if (k < F) {
                    P[(size_t)b * F + k] = pv[m];
                    g[k] = 1.0 / sqrt(pv[m]);
                }
            }
            __syncthreads();
        }
        irfft_merge_conj(z, N, tw, g, s_yN);
        smem_fft<LGN>(z, tws);
        const double sc = 1.0 / (double)N;
        for (int n = threadIdx.x; n < N; n += 256) {
            const double2 v = z[zp(n)];
            *reinterpret_cast<float2*>(dst + 2 * n) = make_float2((float)(v.x * sc), (float)(-v.y * sc));
        }
        __syncthreads();
    }
}

// gwf_apply_psd in one kernel: rfft(sig) * gain(P) -> irfft, gain per `mode` as in spectral_scale_kernel (no mean removal)
template <int LGN>
__global__ void __launch_bounds__(256, 2) apply_psd_fused_kernel(const float* __restrict__ sig, const double* __restrict__ P,
                                                                 long p_b_stride, int mode, const double2* __restrict__ tw,
                                                                 float* __restrict__ o32, double* __restrict__ o64) {
    extern __shared__ __align__(16) unsigned char wf_smem[];
    __shared__ double s_yN;
    constexpr int N = 1 << LGN, L = 2 * N, F = N + 1;
    const int b = blockIdx.x;
    double2* z = reinterpret_cast<double2*>(wf_smem);
    double2* tws = z + (N + N / 6 + 8);
    double* g = reinterpret_cast<double*>(tws + (N / 2 + N / 12 + 8));
    for (int j = threadIdx.x; j < N / 2; j += 256) tws[tw_skew(j)] = tw[2 * j];
    const float* src = sig + (size_t)b * L;
    for (int n = threadIdx.x; n < N; n += 256) {
        const float2 v = *reinterpret_cast<const float2*>(src + 2 * n);
        z[zp(n)] = make_double2((double)v.x, (double)v.y);
    }
    for (int k = threadIdx.x; k < F; k += 256) {
        const double p = P[(size_t)b * p_b_stride + k];
        g[k] = mode == 2 ? sqrt(p + 1e-12) : 1.0 / sqrt(mode == 3 ? p + 1e-20 : p + 1e-12);
    }
    __syncthreads();
    smem_fft<LGN>(z, tws);
    rfft_split(z, N, tw, &s_yN);
    irfft_merge_conj(z, N, tw, g, s_yN);
    smem_fft<LGN>(z, tws);
    const double sc = 1.0 / (double)N;
    for (int n = threadIdx.x; n < N; n += 256) {
        const double2 v = z[zp(n)];
        const double a = v.x * sc, c = -v.y * sc;
        if (o32 != nullptr) *reinterpret_cast<float2*>(o32 + (size_t)b * L + 2 * n) = make_float2((float)a, (float)c);
        if (o64 != nullptr) *reinterpret_cast<double2*>(o64 + (size_t)b * L + 2 * n) = make_double2(a, c);
    }
}

int g_whiten_fused = 1;       // gwf_set_option("fused", 0): always take the cuFFT path
extern "C" int gwf_set_option(const char* name, int value) {
    if (strcmp(name, "fused") == 0) { g_whiten_fused = value; return GW_OK; }
    gw_set_error("gwf_set_option: unknown option %s", name);
    return GW_ERR_ARG;
}

// _whiten_pair_train_like (inference.py:137-153): y, x fp32 [B, L] (x may be NULL) -> y_w, x_w fp32 [B, L], P fp64 [B, L/2+1]
extern "C" int gwf_whiten_train_like(const float* y, const float* x, int B, int L, float* y_w, float* x_w, double* P, void* work,
                                     void* stream) {
    GW_REQUIRE(B > 0 && L >= 2 && y && y_w && P && work, "gwf_whiten_train_like: arguments");
    GW_REQUIRE((x == nullptr) == (x_w == nullptr), "gwf_whiten_train_like: x / x_w mismatch");
    cudaStream_t st = (cudaStream_t)stream;
    const int F = L / 2 + 1;
    if (g_whiten_fused && L >= 64 && L <= 8192 && (L & (L - 1)) == 0) {
        int lgN = 0;
        while ((2 << lgN) < L) ++lgN;
        const double2* tw = nullptr;
        int rct = get_twiddles(L, st, &tw);
        if (rct != GW_OK) return rct;
        const size_t smem = (size_t)(L / 2 + L / 12 + 8 + L / 4 + L / 24 + 8) * sizeof(double2) + (size_t)F * sizeof(double) + 16;
#define WF_GO(LG)                                                                                                              \
    case LG:                                                                                                                   \
        if (smem > 48 * 1024)                                                                                                  \
            GW_CUDA(cudaFuncSetAttribute(whiten_fused_kernel<LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        whiten_fused_kernel<LG><<<B, 256, smem, st>>>(y, x, tw, y_w, x_w, P);                                                  \
        break;
        switch (lgN) {
            WF_GO(5) WF_GO(6) WF_GO(7) WF_GO(8) WF_GO(9) WF_GO(10) WF_GO(11) WF_GO(12)
            default: GW_REQUIRE(false, "gwf_whiten_train_like: fused length %d", L);
        }
#undef WF_GO
        GW_LAUNCH_CHECK();
        return GW_OK;
    }
    double* tbuf = (double*)work;
    cufftDoubleComplex* Y = (cufftDoubleComplex*)(tbuf + (size_t)B * L);
    cufftDoubleComplex* Z = Y + (size_t)B * F;
    int rc = fft_forward(y, 1, B, L, tbuf, Y, st);
    if (rc != GW_OK) return rc;
    dim3 g(gw_cdiv(F, 256), B);
    periodogram_kernel<<<g, 256, 0, st>>>(Y, F, P);
    GW_LAUNCH_CHECK();
    spectral_scale_kernel<<<g, 256, 0, st>>>(Y, P, F, F, 0, Z);
    GW_LAUNCH_CHECK();
    if ((rc = fft_inverse(Z, B, L, tbuf, y_w, nullptr, st)) != GW_OK) return rc;
    if (x != nullptr) {
        if ((rc = fft_forward(x, 1, B, L, tbuf, Y, st)) != GW_OK) return rc;
        spectral_scale_kernel<<<g, 256, 0, st>>>(Y, P, F, F, 0, Z);
        GW_LAUNCH_CHECK();
        if ((rc = fft_inverse(Z, B, L, tbuf, x_w, nullptr, st)) != GW_OK) return rc;
    }
    return GW_OK;
}

// spectral multiply / divide by a given PSD: mode 1 = whiten with 1/sqrt(P + 1e-12) (_whiten_pair_model, inference.py:190-199, no
// mean removal), mode 2 = de-whiten with sqrt(P + 1e-12) (_dewhiten_train_like / _dewhiten_model, inference.py:155-159, 201-203),
// mode 3 = whiten with 1/sqrt(P + 1e-20) (the data loader's model / Welch whitening, dataloader.py:127-143).
// P fp64 [B, L/2+1] or one shared row (p_shared != 0).  Output fp32 (o32) and / or fp64 (o64).
extern "C" int gwf_apply_psd(const float* sig, int B, int L, const double* P, int p_shared, int mode, float* o32, double* o64,
                             void* work, void* stream) {
    GW_REQUIRE(B > 0 && L >= 2 && sig && P && work && (o32 || o64) && (mode >= 1 && mode <= 3), "gwf_apply_psd: arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int F = L / 2 + 1;
    if (g_whiten_fused && L >= 64 && L <= 8192 && (L & (L - 1)) == 0) {
        int lgN = 0;
        while ((2 << lgN) < L) ++lgN;
        const double2* tw = nullptr;
        int rct = get_twiddles(L, st, &tw);
        if (rct != GW_OK) return rct;
        const size_t smem = (size_t)(L / 2 + L / 12 + 8 + L / 4 + L / 24 + 8) * sizeof(double2) + (size_t)F * sizeof(double) + 16;
        const long pbs = p_shared ? 0 : F;
#define AP_GO(LG)                                                                                                              \
    case LG:                                                                                                                   \
        if (smem > 48 * 1024)                                                                                                  \
            GW_CUDA(cudaFuncSetAttribute(apply_psd_fused_kernel<LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        apply_psd_fused_kernel<LG><<<B, 256, smem, st>>>(sig, P, pbs, mode, tw, o32, o64);                                     \
        break;
        switch (lgN) {
            AP_GO(5) AP_GO(6) AP_GO(7) AP_GO(8) AP_GO(9) AP_GO(10) AP_GO(11) AP_GO(12)
            default: GW_REQUIRE(false, "gwf_apply_psd: fused length %d", L);
        }
#undef AP_GO
        GW_LAUNCH_CHECK();
        return GW_OK;
    }
    double* tbuf = (double*)work;
    cufftDoubleComplex* Y = (cufftDoubleComplex*)(tbuf + (size_t)B * L);
    cufftDoubleComplex* Z = Y + (size_t)B * F;
    int rc = fft_forward(sig, 0, B, L, tbuf, Y, st);
    if (rc != GW_OK) return rc;
    dim3 g(gw_cdiv(F, 256), B);
    spectral_scale_kernel<<<g, 256, 0, st>>>(Y, P, p_shared ? 0 : F, F, mode, Z);
    GW_LAUNCH_CHECK();
    return fft_inverse(Z, B, L, tbuf, o32, o64, st);
}

// _interp_psd_for_length (inference.py:181-188): np.interp of a model PSD given on rfftfreq(2*(n_src-1), 1/fs) onto rfftfreq(L, 1/fs)
__global__ void interp_psd_kernel(const double* __restrict__ Ps, int n_src, int L, double fs, double* __restrict__ out) {
    const int F = L / 2 + 1;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    if (n_src == F) { out[f] = Ps[f]; return; }
    const int Ls = n_src * 2 - 2;
    const double df_s = 1.0 / ((double)Ls * (1.0 / fs)), df_t = 1.0 / ((double)L * (1.0 / fs));   // rfftfreq: k / (n*d)
    const double x = (double)f * df_t;
    // np.interp: left/right = end values; linear in between on the grid xp[k] = k * df_s
    if (x <= 0.0) { out[f] = Ps[0]; return; }
    const double xmax = (double)(n_src - 1) * df_s;
    if (x >= xmax) { out[f] = Ps[n_src - 1]; return; }
    int k = (int)(x / df_s);
    if (k > n_src - 2) k = n_src - 2;
    while (k > 0 && (double)k * df_s > x) --k;
    while (k < n_src - 2 && (double)(k + 1) * df_s <= x) ++k;
    const double x0 = (double)k * df_s, x1 = (double)(k + 1) * df_s;
    const double slope = (Ps[k + 1] - Ps[k]) / (x1 - x0);
    out[f] = slope * (x - x0) + Ps[k];
}
extern "C" int gwf_interp_psd(const double* P_src, int n_src, int L, double fs, double* out, void* stream) {
    GW_REQUIRE(P_src && out && n_src >= 2 && L >= 2 && fs > 0.0, "gwf_interp_psd: arguments");
    interp_psd_kernel<<<gw_cdiv(L / 2 + 1, 256), 256, 0, (cudaStream_t)stream>>>(P_src, n_src, L, fs, out);
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// batched form: P_src fp64 [B, n_src] -> out [B, L/2+1]
__global__ void interp_psd_batch_kernel(const double* __restrict__ Ps, int n_src, int L, double fs, double* __restrict__ out) {
    const int F = L / 2 + 1;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const double* ps = Ps + (size_t)blockIdx.y * n_src;
    double* o = out + (size_t)blockIdx.y * F;
    if (n_src == F) { o[f] = ps[f]; return; }
    const int Ls = n_src * 2 - 2;
    const double df_s = 1.0 / ((double)Ls * (1.0 / fs)), df_t = 1.0 / ((double)L * (1.0 / fs));
    const double x = (double)f * df_t;
    if (x <= 0.0) { o[f] = ps[0]; return; }
    const double xmax = (double)(n_src - 1) * df_s;
    if (x >= xmax) { o[f] = ps[n_src - 1]; return; }
    int k = (int)(x / df_s);
    if (k > n_src - 2) k = n_src - 2;
    while (k > 0 && (double)k * df_s > x) --k;
    while (k < n_src - 2 && (double)(k + 1) * df_s <= x) ++k;
    const double x0 = (double)k * df_s, x1 = (double)(k + 1) * df_s;
    o[f] = (ps[k + 1] - ps[k]) / (x1 - x0) * (x - x0) + ps[k];
}
extern "C" int gwf_interp_psd_batch(const double* P_src, int B, int n_src, int L, double fs, double* out, void* stream) {
    GW_REQUIRE(P_src && out && B > 0 && n_src >= 2 && L >= 2 && fs > 0.0, "gwf_interp_psd_batch: arguments");
    interp_psd_batch_kernel<<<dim3(gw_cdiv(L / 2 + 1, 256), B), 256, 0, (cudaStream_t)stream>>>(P_src, n_src, L, fs, out);
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// np.interp(rfftfreq(L, 1/fs), xp, fp, left = fp[0], right = fp[-1]) for an arbitrary increasing grid xp (a saved Welch PSD with
// its own frequency array, dataloader.py:136-139): xp, fp fp64 [B, n_src] -> out [B, L/2+1]
__global__ void interp_grid_kernel(const double* __restrict__ xp, const double* __restrict__ fp, int n_src, int L, double fs,
                                   double* __restrict__ out) {
    const int F = L / 2 + 1;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const double* xs = xp + (size_t)blockIdx.y * n_src;
    const double* ps = fp + (size_t)blockIdx.y * n_src;
    const double x = (double)f / ((double)L * (1.0 / fs));
    double v;
    if (x <= xs[0]) v = ps[0];
    else if (x >= xs[n_src - 1]) v = ps[n_src - 1];
    else {
        int lo = 0, hi = n_src - 1;                 // xs[lo] <= x < xs[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (xs[mid] <= x) lo = mid; else hi = mid;
        }
        v = (ps[lo + 1] - ps[lo]) / (xs[lo + 1] - xs[lo]) * (x - xs[lo]) + ps[lo];
    }
    out[(size_t)blockIdx.y * F + f] = v;
}
extern "C" int gwf_interp_grid(const double* xp, const double* fp, int B, int n_src, int L, double fs, double* out, void* stream) {
    GW_REQUIRE(xp && fp && out && B > 0 && n_src >= 2 && L >= 2 && fs > 0.0, "gwf_interp_grid: arguments");
    interp_grid_kernel<<<dim3(gw_cdiv(L / 2 + 1, 256), B), 256, 0, (cudaStream_t)stream>>>(xp, fp, n_src, L, fs, out);
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ---------------------------------------------------------------------------------------------- Welch PSD (scipy.signal.welch)
// inference._whiten_pair_welch (inference.py:161-173): welch(y, fs, nperseg = min(4096, L)) with scipy's defaults -- periodic Hann
// window, noverlap = nperseg // 2, constant detrend per segment, one-sided density scaling, mean over the segments.
// segment (b, s): (y[s*step + n] - mean_n) * w[n] -> fp64 [B * n_seg, nperseg]
__global__ void __launch_bounds__(256) welch_segment_kernel(const float* __restrict__ y, int L, int nperseg, int step, int n_seg,
                                                            double* __restrict__ out) {
    __shared__ double red[8];
    const int b = blockIdx.x / n_seg, sgm = blockIdx.x % n_seg;
    const float* ys = y + (size_t)b * L + (size_t)sgm * step;
    double s = 0.0;
    for (int i = threadIdx.x; i < nperseg; i += 256) s += (double)ys[i];
    const double mean = blk_sum_d(s, red) / (double)nperseg;
    double* o = out + (size_t)blockIdx.x * nperseg;
    for (int i = threadIdx.x; i < nperseg; i += 256) {
        const double w = 0.5 - 0.5 * cos(2.0 * 3.14159265358979323846 * (double)i / (double)nperseg);
        o[i] = ((double)ys[i] - mean) * w;
    }
}
// Pxx[b, f] = mean_s |Y[b, s, f]|^2 * scale * (2 for the bins that have a mirror image)
__global__ void __launch_bounds__(256) welch_average_kernel(const cufftDoubleComplex* __restrict__ Y, int n_seg, int nperseg, double scale,
                                                            double* __restrict__ Pxx) {
    const int F = nperseg / 2 + 1;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int b = blockIdx.y;
    double acc = 0.0;
    for (int sgm = 0; sgm < n_seg; ++sgm) {
        const cufftDoubleComplex v = Y[((size_t)b * n_seg + sgm) * F + f];
        acc += v.x * v.x + v.y * v.y;
    }
    const bool mirrored = f > 0 && ((nperseg & 1) || f < F - 1);
    Pxx[(size_t)b * F + f] = acc / (double)n_seg * scale * (mirrored ? 2.0 : 1.0);
}
static int welch_nseg(int L, int nperseg) {
    const int nov = nperseg / 2, step = nperseg - nov;
    return (L - nov) / step;
}
extern "C" long gwf_welch_workspace_bytes(int B, int L, int nperseg) {
    if (nperseg < 2 || nperseg > L) return 0;
    const long ns = welch_nseg(L, nperseg);
    return (long)B * ns * nperseg * 8 + (long)B * ns * (nperseg / 2 + 1) * 16;
}
extern "C" int gwf_welch_psd(const float* y, int B, int L, double fs, int nperseg, double* Pxx, void* work, void* stream) {
    GW_REQUIRE(y && Pxx && work && B > 0 && nperseg >= 2 && nperseg <= L && fs > 0.0, "gwf_welch_psd: arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_seg = welch_nseg(L, nperseg), step = nperseg - nperseg / 2, F = nperseg / 2 + 1;
    GW_REQUIRE(n_seg >= 1, "gwf_welch_psd: no complete segment");
    double* tbuf = (double*)work;
    cufftDoubleComplex* Y = (cufftDoubleComplex*)(tbuf + (size_t)B * n_seg * nperseg);
    welch_segment_kernel<<<B * n_seg, 256, 0, st>>>(y, L, nperseg, step, n_seg, tbuf);
    GW_LAUNCH_CHECK();
    cufftHandle h;
    int rc = get_plan(0, nperseg, B * n_seg, &h);
    if (rc != GW_OK) return rc;
    GW_CUFFT(cufftSetStream(h, st));
    GW_CUFFT(cufftExecD2Z(h, tbuf, Y));
    double w2 = 0.0;                                        // sum of the squared window (host: nperseg <= a few thousand terms)
    for (int i = 0; i < nperseg; ++i) {
        const double w = 0.5 - 0.5 * cos(2.0 * 3.14159265358979323846 * (double)i / (double)nperseg);
        w2 += w * w;
    }
    welch_average_kernel<<<dim3(gw_cdiv(F, 256), B), 256, 0, st>>>(Y, n_seg, nperseg, 1.0 / (fs * w2), Pxx);
    GW_LAUNCH_CHECK();
    return GW_OK;
}

// ---------------------------------------------------------------------------------------------- sigma (_pick_sigma)
// mode 0: np.std(y64) (population);  mode 1: 1.4826 * median(|y64 - median(y64)|) + 1e-24  (bitonic sort in shared memory)
__device__ void bitonic_sort(double* v, int n) {        // n = power of two, all threads of the CTA participate
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const bool up = (i & k) == 0;
                    const double a = v[i], c = v[p];
                    if ((a > c) == up) { v[i] = c; v[p] = a; }
                }
            }
            __syncthreads();
        }
}
__device__ double median_sorted(const double* v, int L) { return (L & 1) ? v[L / 2] : 0.5 * (v[L / 2 - 1] + v[L / 2]); }

__global__ void __launch_bounds__(256) sigma_kernel(const float* __restrict__ y, int L, int n2, int mode, double* __restrict__ out) {
    extern __shared__ double sv[];
    __shared__ double red[8];
    const int b = blockIdx.x;
    const float* yb = y + (size_t)b * L;
    if (mode == 0) {
        double s = 0.0;
        for (int i = threadIdx.x; i < L; i += 256) s += (double)yb[i];
        const double mean = blk_sum_d(s, red) / (double)L;
        double q = 0.0;
        for (int i = threadIdx.x; i < L; i += 256) {
            const double d = (double)yb[i] - mean;
            q += d * d;
        }
        q = blk_sum_d(q, red);
        if (threadIdx.x == 0) out[b] = sqrt(q / (double)L);
        return;
    }
    for (int i = threadIdx.x; i < n2; i += 256) sv[i] = i < L ? (double)yb[i] : 1.0e300;
    __syncthreads();
    bitonic_sort(sv, n2);
    const double med = median_sorted(sv, L);
    __syncthreads();
    for (int i = threadIdx.x; i < n2; i += 256) sv[i] = i < L ? fabs((double)yb[i] - med) : 1.0e300;
    __syncthreads();
    bitonic_sort(sv, n2);
    if (threadIdx.x == 0) out[b] = 1.4826 * median_sorted(sv, L) + 1e-24;
}
extern "C" int gwf_sigma(const float* y, int B, int L, int mode, double* out, void* stream) {
    GW_REQUIRE(y && out && B > 0 && L > 0 && (mode == 0 || mode == 1), "gwf_sigma: arguments");
    int n2 = 1;
    while (n2 < L) n2 <<= 1;
    const size_t smem = mode == 1 ? (size_t)n2 * sizeof(double) : 0;
    GW_REQUIRE(smem <= 200 * 1024, "gwf_sigma: MAD needs L <= 16384 (L=%d)", L);
    if (smem > 48 * 1024) GW_CUDA(cudaFuncSetAttribute(sigma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sigma_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(y, L, n2, mode, out);
    GW_LAUNCH_CHECK();
    return GW_OK;
}
