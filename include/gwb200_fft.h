/*
 * gwb200_fft -- C-ABI of the GPU whitening / de-whitening / sigma kernels around the reverse chain.
 * Separate library (libgwb200_fft.so, links cuFFT).  Replaces, batched and in fp64, the numpy helpers of the reference:
 *   gwf_whiten_train_like  <- inference._whiten_pair_train_like   (inference.py:137-153; dataloader.py:110-151)
 *   gwf_apply_psd (mode 1) <- inference._whiten_pair_model         (inference.py:190-199), after gwf_interp_psd
 *   gwf_apply_psd (mode 2) <- inference._dewhiten_train_like / _dewhiten_model / _dewhiten_welch (inference.py:155-159, 175-179, 201-203)
 *   gwf_apply_psd (mode 3) <- dataloader._whiten_with_model_psd / _whiten_with_welch (1e-20 floor; dataloader.py:127-143)
 *   gwf_welch_psd          <- scipy.signal.welch inside inference._whiten_pair_welch (inference.py:161-173)
 *   gwf_interp_psd         <- inference._interp_psd_for_length     (inference.py:181-188)
 *   gwf_sigma              <- inference._pick_sigma / _mad_std     (inference.py:36-38, 125-135)
 * Conventions as in gwb200.h (device pointers, void* stream, 0 / negative status, gwf_last_error()).  cuFFT plans are created
 * on first use per (L, B) and cached, so the first call of a shape allocates; `work` >= gwf_workspace_bytes(B, L) bytes.
 */
#ifndef GWB200_FFT_H
#define GWB200_FFT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
const char* gwf_last_error(void);
/* A/B switches: "fused" (default 1): power-of-two L <= 8192 whitens in ONE kernel per call (shared-memory fp64 FFT); 0 = cuFFT path */
int gwf_set_option(const char* name, int value);
long gwf_workspace_bytes(int B, int L);
int gwf_whiten_train_like(const float* y, const float* x, int B, int L, float* y_w, float* x_w, double* P, void* work,
                          void* stream);
int gwf_apply_psd(const float* sig, int B, int L, const double* P, int p_shared, int mode, float* o32, double* o64, void* work,
                  void* stream);
int gwf_interp_psd(const double* P_src, int n_src, int L, double fs, double* out, void* stream);
/* batched: P_src fp64 [B, n_src] on rfftfreq(2 (n_src - 1), 1/fs) -> out [B, L/2+1] */
int gwf_interp_psd_batch(const double* P_src, int B, int n_src, int L, double fs, double* out, void* stream);
/* np.interp(rfftfreq(L, 1/fs), xp, fp) with end values outside, xp / fp fp64 [B, n_src] (a saved Welch PSD with its own frequency
 * array: dataloader.py:136-139) -> out [B, L/2+1] */
int gwf_interp_grid(const double* xp, const double* fp, int B, int n_src, int L, double fs, double* out, void* stream);
/* scipy.signal.welch(y, fs, nperseg) with scipy's defaults (periodic Hann, 50 % overlap, constant detrend, one-sided density,
 * mean over segments) <- inference._whiten_pair_welch (inference.py:161-173).  y fp32 [B, L] -> Pxx fp64 [B, nperseg/2+1];
 * work >= gwf_welch_workspace_bytes(B, L, nperseg) bytes. */
long gwf_welch_workspace_bytes(int B, int L, int nperseg);
int gwf_welch_psd(const float* y, int B, int L, double fs, int nperseg, double* Pxx, void* work, void* stream);
int gwf_sigma(const float* y, int B, int L, int mode, double* out, void* stream);
#ifdef __cplusplus
}
#endif
#endif
