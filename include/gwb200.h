/*
 * gwb200 -- C-ABI of the B200 (sm_100a) kernels behind the snr_denoising hot path.
 *
 * The reference (Ch4rlesSm1th99/Diffusion_Models_for_Gravitational_Waveform_Reconstruction) has no
 * FFI of its own: its hot path is Python calling torch ATen ops.  Each entry point below replaces a
 * group of those call sites (cited as file:line under src/snr_denoising/).  The Python mirror of the
 * reference API (package diffusion_models_for_gravitational_waveform_reconstruction_b200) binds these
 * through ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch), except where noted "host";
 *   - `stream` is a cudaStream_t passed as void*; nothing allocates, synchronises or reads back, so
 *     every call is CUDA-graph capturable;
 *   - return 0 on success, negative on error (gw_last_error() gives the text);
 *   - activations are channels-last [B, L, C] in `dtype` storage (GW_F32 = 0, GW_BF16 = 1), C % 64 == 0;
 *     network inputs / outputs keep the reference's [B, C, L] fp32 layout;
 *   - GroupNorm always has 8 groups (gcd(8, C) with C % 64 == 0; models.py:154-158).
 */
#ifndef GWB200_H
#define GWB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GW_F32 0
#define GW_BF16 1
#define GW_DOTS 2 /* gw_final_step only: `h` holds the head dot products [Bn, L, 4] fp32 written by gw_conv_gn2 */
#define GW_MAX_LEVELS 8

int gw_version(void);
/* runtime switches for A/B measurements: "gn_bwd_stream" (1 = HBM-streaming GroupNorm backward kernels, default),
 * "gn_bwd_stats_fast" (1 = their compile-time-specialised versions + analytic conv-bias gradient, default),
 * "gn_bwd_fused" / "gn_bwd_fused_slice" (one-pass GroupNorm backward of gw_gn_bwd2: enable, largest shared-memory slice in bytes),
 * "final_stream" (1 = HBM-streaming head + update kernel for bf16 / C = 64, default),
 * "pdl" (1 = kernels are launched with programmatic stream serialization: a kernel's launch overlaps the previous kernel's tail,
 *  every kernel executes griddepcontrol.wait before it touches memory; default 0 -- measured without gain inside CUDA graphs),
 * "pair2" (1 = CTA-pair flavour of the fused conv block kernel, tcgen05 cta_group::2; default 0),
 * "one_group" (1 = launches with at most one sample per CTA group use the single-group epilogue flavour; default 1; results are
 *  bit-identical either way), "conv_in_mma" / "wgrad_in_mma" / "final_bwd_stream" (first-layer and head kernels: tensor-core /
 *  streaming versions, default 1) */
int gw_set_option(const char* name, int value);
const char* gw_last_error(void);
int gw_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- time conditioning: TimeEmbedding + time_mlp + all tproj_* (models.py:19-31, 105-109, 137-142, 197).
 * t: int64 [n]; w1 [base, time_dim], b1 [base]; w2 [F, base], b2 [F] = the 2*depth+1 tproj Linear layers
 * concatenated in order enc0..encD-1, mid, dec0..decD-1; out fp32 [n, F] rows of (gamma|beta) blocks.
 * aux (optional, training): fp32 [n, time_dim + 3*base] rows of [emb | pre | ctx | act] kept for gw_film_bwd. */
int gw_film_vectors(const int64_t* t, int n, int time_dim, float max_time, const float* w1, const float* b1,
                    const float* w2, const float* b2, int base, int F, float* out, float* aux, void* stream);

/* ---- conditioning pyramid: F.interpolate(cond, size=L_j, mode="linear", align_corners=False) for every
 * level j (models.py:188-193), written channels-last fp32 [B, L_j, Cc].
 * x: fp32 [B, Cx, L] network input; the cond channels are x[:, 1:1+Cc]. */
int gw_cond_pyramid(const float* x, int B, int Cx, int L, int Cc, int n_levels, const int* level_len /*host*/,
                    float* const* level_out /*host array of device ptrs*/, void* stream);

/* ---- first encoder conv: Conv1d(Cx -> C, k=3, pad=1) on the [B, Cx, L] fp32 input (models.py:204, K1) with the
 * GroupNorm partial statistics fused in the epilogue.  w [C, Cx, 3], bias [C]; raw [B, L, C] (dtype);
 * part fp32 [B, n_part, 8, 2] = (sum, sum of squares) per 128-position tile; n_part = ceil(L/128).
 * x_alt / alt_stride: optional ping-pong input selected by (*step_ptr & 1) (see gw_final_step). */
int gw_conv_in(const float* x, const float* x_alt, const int* step_ptr, int B, int Cx, int L, const float* w,
               const float* bias, int C, void* raw, int dtype, float* part, void* stream);

/* ---- fused first block for inference (models.py:204-208 without the raw tensor): pass 1 accumulates the GroupNorm partial
 * sums of the conv output, pass 2 recomputes the conv and applies GroupNorm + SiLU + cond 1x1 conv + FiLM (+ avg_pool).
 * The conditioning channels are x[:, 1:1+Cc] themselves (level-0 interpolation is the identity).  film row selection as in
 * gw_gn_apply.  out [B, L, C], pooled [B, L/2, C] or NULL (dtype); part: scratch fp32 [B, ceil(L/128), 8, 2]. */
int gw_conv_in_block(const float* x, const float* x_alt, const int* step_ptr, int B, int Cx, int L, const float* w,
                     const float* bias, int C, const float* gn_w, const float* gn_b, int Cc, const float* wc,
                     const float* bc, const float* film, int film_off, long film_b_stride, long film_step_stride,
                     void* out, void* pooled, int dtype, float* part, void* stream);

/* ---- generic Conv1d(k=3, pad=1) on channels-last activations, CUDA-core fp32 math (exact mode, any L).
 * Input is the virtual concat [nearest-upsample x2 (src0) | src1] (models.py:217-222); src1 may be NULL and
 * `up0` = 0 for encoder/mid convs.  src0 [B, L0, C0], src1 [B, L, C1]; w3 = the reference weight [Cout, C0+C1, 3] fp32;
 * raw [B, L, Cout]; part [B, ceil(L/64), 8, 2].  bias and part may be NULL (dgrad use: see gw_weight_dgrad). */
int gw_conv3_simt(const void* src0, int C0, int L0, int up0, const void* src1, int C1, int B, int L,
                  const float* w3, const float* bias, int Cout, void* raw, int dtype, float* part, void* stream);

/* ---- fused GroupNorm-apply + SiLU + conditioning 1x1 conv + FiLM (+ skip write, + avg_pool1d(2,2) write)
 * (models.py:165-166, 188-193, 169-173, 207-208; K9-K13).
 * raw [B, L, C]; part [B, n_part, 8, 2]; gn_w, gn_b [C]; cond fp32 [B, L, Cc] or NULL; wc [C, Cc], bc [C];
 * film: fp32 row(s) of gw_film_vectors output, this layer's block at `film_off`; row for sample b is
 *   film + ((step_ptr ? *step_ptr : 0) * film_step_stride + b * film_b_stride) + film_off;
 * out [B, L, C]; pooled [B, L/2, C] or NULL; stats_out fp32 [B, 8, 2] (mean, rstd) or NULL. */
int gw_gn_apply(const void* raw, const float* part, int n_part, int B, int L, int C, const float* gn_w,
                const float* gn_b, const float* cond, int Cc, const float* wc, const float* bc, const float* film,
                int film_off, long film_b_stride, long film_step_stride, const int* step_ptr, void* out,
                void* pooled, float* stats_out, int dtype, void* stream);

/* ---- the same operation as gw_gn_apply for bf16 tensors, HBM-streaming implementation (stream_gn.cu): contiguous row
 * ranges go through a ring of shared-memory stages with 1-D bulk copies (cp.async.bulk) in and bulk stores out.
 * Needs L % 4 == 0 and C in {64, 128, 256}. */
int gw_gn_apply_stream(const void* raw, const float* part, int n_part, int B, int L, int C, const float* gn_w,
                       const float* gn_b, const float* cond, int Cc, const float* wc, const float* bc, const float* film,
                       int film_off, long film_b_stride, long film_step_stride, const int* step_ptr, void* out,
                       void* pooled, float* stats_out, void* stream);

/* ---- FIRST block for inference in one pass with analytic GroupNorm statistics (conv_in_direct.cu; models.py:160-173,
 * 188-193, 204-208).  The first conv has only 3*Cx <= 24 inputs per output, so the first / second moments of its output are a
 * quadratic form of the input's lag-(0,1,2) cross products: one small kernel reads the fp32 input once and leaves the
 * per-(sample, channel) epilogue coefficients in coef_ws (gw_conv_in_direct_ws_floats floats), then any CTA convolves any
 * 256-row slice (tf32 mma.sync) and applies GroupNorm / SiLU / cond / FiLM / pool on the accumulator fragments -- no exchange
 * between CTAs, no raw tensor.  Same arguments as gw_conv_in_gn; x / x_alt / step_ptr: ping-pong input of the sampler.
 * Needs Cx <= 8, C == 64, L even. */
long gw_conv_in_direct_ws_floats(int B, int Cx, int L, int C, int Cc);
int gw_conv_in_direct(const float* x, const float* x_alt, const int* step_ptr, int B, int Cx, int L, const float* w,
                      const float* bias, int C, const float* gn_w, const float* gn_b, int Cc, const float* wc,
                      const float* bc, const float* film, int film_off, long film_b_stride, long film_step_stride,
                      void* out, void* pooled, float* coef_ws, void* stream);

/* ---- head: final Conv1d(C+1 -> 1, k=3) on cat[h, x_t] (models.py:227-230, K8) optionally fused with classifier-free
 * guidance combine and the DDIM/DDPM update (inference.py:445-484, K20/K21).
 *
 * h [Bn, L, C] (dtype); net_in fp32 [Bn, Cx, L] holds x_t in channel 0, y in channel 1, self-cond in channel Cx-1.
 * mode 0 (forward only): eps_out[Bn, L] = conv.
 * mode 1 (sampler step): Bn = B * (cfg_both ? 2 : 1); rows [0,B) are the conditional batch, [B,2B) the
 *   unconditional one.  Per step s = *step_ptr the kernel reads coef[s*16 + ..]:
 *     0 sqrt(1-ab_t)  1 sqrt(ab_t)  2 sqrt(ab_prev)  3 sqrt(max(1-ab_prev-sigma^2,0))  4 sigma  5 w_cfg
 *     6 use (0 cond only, 1 uncond only, 2 both)  7 last (t == 0)  8 noise draw index (as float)  9 sqrt(max(1-ab_t,1e-12))
 *   and writes x_{t-1} to channel 0 and x0_hat to channel Cx-1 (if selfcond) of the OTHER ping-pong buffer
 *   (net_out) for every row of Bn.
 *   noise: fp32 [n_draws, B, L] injected draws or NULL -> Philox(seed, sample0 + b, step).
 *   eps_out / x0_out (fp32 [B, L]) are optional traces. */
typedef struct {
    int mode;          /* 0 forward, 1 step */
    int cfg_both;      /* Bn = 2B */
    int selfcond;      /* write x0_hat to channel Cx-1 */
    int pred_x0;       /* pred_type == "x0" */
    float eps_scale;
    float dc_weight;
    const float* y_dc; /* device fp32 [B, L]: the unscaled y for the dc blend (inference.py:472); NULL if dc_weight == 0 */
    unsigned long long seed;
    long sample0;      /* global index of sample 0 (Philox stream id, world-size independent) */
    void* advance;     /* dtype = GW_DOTS, mode 1: device uint32 (zeroed once): the last CTA of the launch does *step_ptr += 1, so
                          no gw_step_advance launch is needed between reverse steps; NULL: the caller advances the counter */
    const unsigned long long* rng; /* device uint64[2] = {seed, sample0} read at run time (a captured CUDA graph then follows
                          later changes of the Philox key); NULL: the by-value seed / sample0 above are used */
} gw_step_params;

int gw_final_step(const void* h, int dtype, const float* net_a, const float* net_b, int B, int Cx, int L, int C,
                  const float* wf, const float* bf, const gw_step_params* p /*host*/, const float* coef,
                  const int* step_ptr, const float* noise, float* eps_out, float* x0_out, void* stream);

/* ---- step counter for graph replay: *step_ptr += 1 (or = value when set >= 0). */
int gw_step_advance(int* step_ptr, int set_value, void* stream);

/* ---- standard normals of the Philox stream (seed, sample0 + b, step) -> out fp32 [B, L].  The sampler's x_T
 * (torch.randn at inference.py:409-415) is step 0 of each sample's stream; step s+1 is the noise of reverse step s. */
int gw_philox_normal(unsigned long long seed, long sample0, unsigned step, int B, int L, float* out, void* stream);

/* ---- q_sample (models.py:52-59, K19) fused with the clamp of train.py:381-382 and with network-input packing:
 * x_t = clamp(sqrt(ab[t]) * x0 + sqrt(1-ab[t]) * eps).  x0 fp32 [B, L]; t int64 [B]; eps fp32 [B, L] is READ when
 * philox == 0 and WRITTEN (generated) when philox != 0; x_t goes to net[b, 0, :] (batch stride Cx*L). */
int gw_q_sample(const float* x0, const int64_t* t, const float* sqrt_ab, const float* sqrt_1mab, float* eps,
                int philox, unsigned long long seed, long sample0, unsigned step, float clamp, float* net, int B,
                int Cx, int L, void* stream);

/* ---- tcgen05 / TMA implicit-GEMM Conv1d(k=3) on bf16 channels-last activations (K2-K7).  See conv_tc.cu. */
typedef struct {
    int n_src;               /* 1 or 2 */
    int pair;                /* 1: decoder conv evaluated in pair space (nearest-upsample folded into the weights);
                              * 2: dgrad through the nearest upsample: src0 = d_raw [B, L0 = 2L, C0], output d_h [B, L, Cout]
                              *    = sum of the gradients of the two upsampled positions (weights from gw_weight_dgrad);
                              * 3: plain one-source conv evaluated in pair space (L even, Cout <= 128): same result as pair = 0,
                              *    fills a 128-column MMA tile when Cout = 64 */
    int B, L;                /* output length L (positions) */
    int C0, L0;              /* src0 channels / length (L0 = L/2 when pair) */
    int C1;                  /* src1 channels (skip), 0 if none */
    int Cout;
} gw_conv_tc_shape;

/* packed weight size in bf16 elements for a shape (host helper) */
long gw_conv_tc_packed_elems(const gw_conv_tc_shape* s);
/* pack fp32 reference weights [Cout, Cin, 3] into the kernel's bf16 segment-major layout (device -> device) */
int gw_conv_tc_pack(const gw_conv_tc_shape* s, const float* w, void* packed, void* stream);
/* run: src0/src1/raw bf16; bias fp32 [Cout] or NULL; part fp32 [B, n_part, 8, 2], n_part = gw_conv_tc_n_part(s), or NULL
 * (no GroupNorm statistics: dgrad use) */
int gw_conv_tc_n_part(const gw_conv_tc_shape* s);
int gw_conv_tc(const gw_conv_tc_shape* s, const void* src0, const void* src1, const void* packed, const float* bias,
               void* raw, float* part, int variant, void* stream);

/* ---- fused block: Conv1d(k=3) -> GroupNorm(8) -> SiLU -> + cond 1x1 conv -> FiLM (-> avg_pool1d(2)) in one tcgen05 kernel
 * (models.py:160-173, 188-193, 205-208, 217-225; see conv_gn.cuh).  gw_conv_tc + gw_gn_apply without the raw round trip:
 * a group of G CTAs keeps one sample's conv output in tensor memory and exchanges GroupNorm partial sums through `sync`.
 *   s: pair 0 (encoder / mid conv) or 1 (decoder conv over cat[upsample(src0), src1]); packed from gw_conv_tc_pack;
 *   cond [B, L, Cc] fp32 (gw_cond_pyramid level) or NULL when Cc == 0; film as in gw_gn_apply;
 *   out [B, L, Cout] bf16; pooled [B, L/2, Cout] bf16 or NULL; raw [B, L, Cout] bf16 or NULL (training keeps the conv output
 *   for gw_gn_bwd); stats_out [B, 8, 2] fp32 (mean, rstd) or NULL;
 *   sync: gw_conv_gn_sync_bytes(B) bytes, zeroed ONCE by the caller, shared by all launches on one stream.
 * gw_conv_gn_group returns G (0: this layer shape / length is not supported -> use gw_conv_tc + gw_gn_apply). */
int gw_conv_gn_group(const gw_conv_tc_shape* s, int Cc, int pool);
long gw_conv_gn_sync_bytes(int B);
int gw_conv_gn(const gw_conv_tc_shape* s, const void* src0, const void* src1, const void* packed, const float* bias,
               const float* gn_w, const float* gn_b, const float* cond, int Cc, const float* wc, const float* bc,
               const float* film, int film_off, long film_b_stride, long film_step_stride, const int* step_ptr,
               void* out, void* pooled, void* raw, float* stats_out, void* sync, void* stream);
/* gw_conv_gn2 = gw_conv_gn that also leaves, for the LAST decoder (Cout = 64 over cat[upsample, skip]), the three dot products
 * of the head conv final(cat[h, x_t]) (models.py:230) per position: head_w = final.weight [C+1, 3], head_dots [B, L, 4] fp32 =
 * (sum_c out[l,c] w[c,0], sum_c out[l,c] w[c,1], sum_c out[l,c] w[c,2], 0), formed from the fp32 epilogue values.  gw_final_step
 * (dtype = GW_DOTS, h = head_dots) finishes eps_hat and the DDIM / DDPM update from 16 B per position.  out may then be NULL
 * (the activated tensor is not written).  head_w == head_dots == NULL: identical to gw_conv_gn. */
int gw_conv_gn2(const gw_conv_tc_shape* s, const void* src0, const void* src1, const void* packed, const float* bias,
                const float* gn_w, const float* gn_b, const float* cond, int Cc, const float* wc, const float* bc,
                const float* film, int film_off, long film_b_stride, long film_step_stride, const int* step_ptr,
                void* out, void* pooled, void* raw, float* stats_out, void* sync, const float* head_w, float* head_dots,
                void* stream);
/* gw_conv_gn3 = gw_conv_gn2 + LAYER CHAINING for the sampler.  serial_ptr (device int32, bumped once per chain by the caller)
 * together with step_ptr gives every (chain, reverse step) a unique value V; a launch with serial_ptr != NULL publishes, per
 * sample, flag[b] = V in its sync buffer once all of the sample's stores have completed.  prev_sync = the sync buffer of the
 * layer that produced src0 in the same reverse step (itself launched with serial_ptr; prev_G = its gw_conv_gn_group): this launch is then issued with
 * programmatic stream serialization, starts on the SMs the producer's early-finishing CTA groups free, and orders itself per
 * SAMPLE through the producer's flags instead of waiting for the producer's whole grid -- no fill / drain bubble between
 * layers, and a partly filled last round overlaps the next layer's work.  Every layer needs its OWN sync buffer. */
int gw_conv_gn3(const gw_conv_tc_shape* s, const void* src0, const void* src1, const void* packed, const float* bias,
                const float* gn_w, const float* gn_b, const float* cond, int Cc, const float* wc, const float* bc,
                const float* film, int film_off, long film_b_stride, long film_step_stride, const int* step_ptr,
                void* out, void* pooled, void* raw, float* stats_out, void* sync_buf, const float* head_w,
                float* head_dots, const void* prev_sync, int prev_G, const int* serial_ptr, void* stream);


/* ---- fused FIRST block for inference: Conv1d(C_in -> 64, k=3) -> GroupNorm -> SiLU -> + cond 1x1 conv -> FiLM (-> pool) in
 * one kernel (models.py:160-173, 188-193, 204-208; conv_in_gn.cu).  gw_conv_in + gw_gn_apply without the raw tensor: G = L/256
 * CTAs share a sample, keep their conv rows in shared memory and exchange GroupNorm sums through `sync` (the buffer of
 * gw_conv_gn).  x / x_alt / step_ptr as in gw_conv_in; the conditioning channels are x[:, 1:1+Cc] themselves.  bf16 outputs.
 * gw_conv_in_gn_group returns G, or 0 when the shape is not supported (C != 64, odd L, L > 16384). */
int gw_conv_in_gn_group(int Cx, int L, int C, int Cc);
int gw_conv_in_gn(const float* x, const float* x_alt, const int* step_ptr, int B, int Cx, int L, const float* w,
                  const float* bias, int C, const float* gn_w, const float* gn_b, int Cc, const float* wc, const float* bc,
                  const float* film, int film_off, long film_b_stride, long film_step_stride, void* out, void* pooled,
                  void* sync, void* stream);

/* =====================================================================================================
 * Training step (train.py:320-456): loss, backward of every block, optimiser.  Parameter gradients are always
 * ACCUMULATED (+=) into fp32 buffers laid out like the reference parameters; the caller zeroes the flat gradient
 * buffer once per step.  `scratch` arguments are caller-owned fp32 device workspaces.
 * ===================================================================================================== */

/* dst[c] (+)= scale * sum_r src[r, c]  -- fixed-order second level of every two-level reduction here.
 * src is scratch: inputs taller than 256 rows are folded in place first. */
int gw_reduce_rows(const float* src, int n_rows, long n_cols, float scale, float* dst, int accumulate, void* stream);

/* masked Huber (loss_type 0, F.smooth_l1_loss beta) / MSE (1) loss of train.py:53-58, 411-421:
 * per_sample[b] = wt[b] * sum_l el*mask / max(sum_l mask, 1); loss[0] = mean_b per_sample;
 * d_eps[b,l] = grad_scale * dloss/d eps_hat[b,l].  mask / wt (= (1-alpha_bar_t)^p, train.py:414-417) may be NULL. */
int gw_loss(const float* eps_hat, const float* eps, const float* mask, const float* wt, int B, int L, int loss_type,
            float beta, float grad_scale, float* per_sample, float* loss, float* d_eps, void* stream);

/* backward of the head conv final(cat[h, x_t]) (models.py:230): d_h [B, L, C] (dtype) or NULL (then the last block's
 * gw_gn_bwd forms it on the fly from do_eps / do_w); d_wf [(C+1)*3], d_bf [1] accumulated.
 * scratch >= B * ceil(L/512) * ((C+1)*3 + 1) floats. */
int gw_final_bwd(const float* d_eps, const void* h, int dtype, const float* net, int B, int Cx, int L, int C,
                 const float* wf, void* d_h, float* scratch, float* d_wf, float* d_bf, void* stream);

/* backward of GroupNorm -> SiLU -> +cond 1x1 conv -> FiLM (-> avg_pool) of one block (models.py:165-173, 188-193, 208).
 * Incoming gradient = do_a [B, L, C] (wrt the block output; NULL if none) + avg_pool backward of do_pool [B, L/2, C]
 * (NULL if none) + (bf16 only) the head-conv gradient sum_k do_w[c,k] * do_eps[b, l-k+1] (do_eps fp32 [B, L], do_w = final.weight).  stats = (mean, rstd) [B, 8, 2] saved by gw_gn_apply.  Writes d_raw [B, L, C] (gradient wrt the conv
 * output), dfilm[b, film_off + (0..C | C..2C)] = (d gamma | d beta); accumulates d_gn_w, d_gn_b, d_bc, d_conv_bias [C],
 * d_wc [C, Cc].  scratch >= gw_gn_bwd_scratch_elems(B, L, C, Cc) floats. */
long gw_gn_bwd_scratch_elems(int B, int L, int C, int Cc);
int gw_gn_bwd(const void* raw, const float* stats, int B, int L, int C, const float* gn_w, const float* gn_b,
              const float* cond, int Cc, const float* wc, const float* bc, const float* film, int film_off,
              long film_b_stride, const void* do_a, const void* do_pool, const float* do_eps, const float* do_w,
              int dtype, float* scratch, float* dfilm, long dfilm_b_stride, void* d_raw, float* d_gn_w,
              float* d_gn_b, float* d_wc, float* d_bc, float* d_conv_bias, void* stream);
/* gw_gn_bwd2 = gw_gn_bwd + `sync` (gw_gn_bwd_sync_bytes(B) bytes, zeroed ONCE by the caller, kept across launches on one
 * stream; may be the gw_conv_gn buffer).  With a non-NULL sync, bf16 layers whose sample fits the shared memory of <= 32 CTAs
 * (gw_gn_bwd_fused_group(L, C, Cc, has_do, has_pool) > 0) run the ONE-PASS kernel (gn_bwd_fused.cu): operands are read once, the
 * CTAs of a sample exchange the GroupNorm group sums through `sync`, d_raw is formed out of shared memory.  Same outputs. */
long gw_gn_bwd_sync_bytes(int B);
int gw_gn_bwd_fused_group(int L, int C, int Cc, int has_do, int has_pool);
int gw_gn_bwd2(const void* raw, const float* stats, int B, int L, int C, const float* gn_w, const float* gn_b,
               const float* cond, int Cc, const float* wc, const float* bc, const float* film, int film_off,
               long film_b_stride, const void* do_a, const void* do_pool, const float* do_eps, const float* do_w,
               int dtype, float* scratch, float* dfilm, long dfilm_b_stride, void* d_raw, float* d_gn_w,
               float* d_gn_b, float* d_wc, float* d_bc, float* d_conv_bias, void* sync, void* stream);

/* gw_gn_bwd in phases, for callers that take the small parameter-gradient reduction off the critical path: phase 0 = gw_gn_bwd;
 * phase 1 = everything d_raw / dfilm need (and, where the parameter kernel is not the last launch, that kernel too); phase 2 = the
 * parameter-gradient kernel alone (same arguments, same scratch, any stream ordered after phase 1; a no-op where phase 1 ran it). */
int gw_gn_bwd_phase(const void* raw, const float* stats, int B, int L, int C, const float* gn_w, const float* gn_b,
                    const float* cond, int Cc, const float* wc, const float* bc, const float* film, int film_off,
                    long film_b_stride, const void* do_a, const void* do_pool, const float* do_eps, const float* do_w,
                    int dtype, float* scratch, float* dfilm, long dfilm_b_stride, void* d_raw, float* d_gn_w,
                    float* d_gn_b, float* d_wc, float* d_bc, float* d_conv_bias, int phase, void* stream);

/* exact-mode conv backward.  gw_weight_dgrad: wt[ci][co][k] = w[co][ci][2-k], so that dgrad = gw_conv3_simt(d_raw, wt).
 * gw_split_cat_grad: gradient of cat[nearest-upsample x2 (h), skip]: d_h [B, L0, C0] (pair sums), d_skip [B, L, C1].
 * gw_wgrad3_simt: dW[co][ci][k] += sum_{b,l} d_raw[b,l,co] * cat[up(src0), src1][b, l+k-1, ci] (split-K partials in scratch).
 * gw_wgrad_in: the same for the first conv, whose input is the fp32 [B, Cx, L] network input. */
int gw_weight_dgrad(const float* w, int Cout, int Cin, float* wt, void* stream);
int gw_split_cat_grad(const void* d_cat, int B, int L, int C0, int L0, int C1, void* d_h, void* d_skip, int dtype,
                      void* stream);
int gw_wgrad3_simt(const void* src0, int C0, int L0, int up0, const void* src1, int C1, const void* d_raw, int B,
                   int L, int Cout, int dtype, float* scratch, long scratch_elems, float* dW, void* stream);
int gw_wgrad_in(const float* x, int B, int Cx, int L, const void* d_raw, int C, int dtype, float* scratch,
                long scratch_elems, float* dW, void* stream);

/* tcgen05 / TMA weight gradient (see wgrad_tc.cu).  mode 0: x [B, L, Cx] is a plain conv input (pooled tensor or skip);
 * mode 1: x [B, L/2, Cx] is h before the nearest upsample.  d_raw [B, L, Cout] bf16.  ACCUMULATES into the input-channel
 * block [ci_off, ci_off+Cx) of dW fp32 [Cout][Cin_total][3].  scratch >= gw_wgrad_tc_scratch_elems(...) floats.
 * variant bit 0: one TMA box per tap instead of row-shifted descriptors; bit 1: split-K by fp32 atomics straight into dW
 * (no fold / scatter passes; summation order, hence the last bits, not reproducible run to run); bit 2: GEMM only -- the
 * deterministic fold + scatter-accumulate into dW is run by gw_wgrad_tc_finish (same shape arguments and scratch), which may be
 * issued on another stream ordered after the GEMM so that the two small kernels leave the critical path. */
long gw_wgrad_tc_scratch_elems(int mode, int B, int L, int Cout, int Cx);
int gw_wgrad_tc(int mode, const void* d_raw, const void* x, int B, int L, int Cout, int Cx, int Cin_total, int ci_off,
                float* scratch, long scratch_elems, float* dW, int variant, void* stream);
int gw_wgrad_tc_finish(int mode, int B, int L, int Cout, int Cx, int Cin_total, int ci_off, float* scratch, float* dW,
                       void* stream);

/* time_mlp / tproj_* backward (models.py:105-109, 137-142): dfilm [B, F] (written by gw_gn_bwd), aux from
 * gw_film_vectors; accumulates dW1 [base, time_dim], db1 [base], dW2 [F, base], db2 [F];
 * scratch >= B*base*(1 + ceil(F/96)) floats (the split-K partials of dfilm x W2). */
int gw_film_bwd(const float* dfilm, const float* aux, const float* w2, int B, int time_dim, int base, int F,
                float* scratch, float* dW1, float* db1, float* dW2, float* db2, void* stream);

/* per-sample training draws, Philox keyed on the global sample index: t ~ U{t_min..T-1} (train.py:376) and the
 * CFG-dropout coin (train.py:386).  step_ptr: device step counter (NULL = 0). */
int gw_train_draws(unsigned long long seed, const int* step_ptr, long sample0, int B, int t_min, int T,
                   float p_uncond, int64_t* t_out, float* drop_out, void* stream);
/* collated batch -> stepper inputs (train.py:336-347, 355-360): clean_out [B0*K, L] = clean_raw / sigma, cond_out [B0*K, 1+Cm, L] =
 * [noisy_raw / sigma | meta], mask_out [B0*K, L] (ones when mask == NULL); row b*K + r <- row b (repeat_interleave of --t_multi).
 * clean_raw, noisy_raw, mask [B0, L]; sigma [B0]; meta [B0, Cm, L] or NULL. */
int gw_batch_prepare(const float* clean_raw, const float* noisy_raw, const float* sigma, const float* mask, const float* meta,
                     int Cm, int B0, int L, int K, float* clean_out, float* cond_out, float* mask_out, void* stream);
/* network-input packing (train.py:350-352, 379-398, 404-407): q_sample with clamp into channel 0, conditioning
 * channels with CFG dropout, zero self-conditioning channel.  clean [B, L]; cond [B, Cc, L]; eps [B, L] read
 * (philox == 0) or generated; drop [B] or NULL. */
int gw_train_pack(const float* clean, const float* cond, int Cc, const int64_t* t, const float* drop,
                  const float* sqrt_ab, const float* sqrt_1mab, float* eps, int philox, unsigned long long seed,
                  long sample0, const int* step_ptr, float clampv, int clamp_y, int drop_all, float* net, int B, int Cx,
                  int L, void* stream);
/* self-conditioning estimate (train.py:40-51): net[b, Cx-1, :] = (x_t - sqrt(1-ab_t) eps_hat) / sqrt(ab_t) */
int gw_selfcond_x0(float* net, const float* eps_hat, const int64_t* t, const float* alpha_bar, int B, int Cx, int L,
                   void* stream);

/* clip_grad_norm_ + AdamW + EMA (train.py:445-455, 73-81) over flat fp32 buffers of n elements.
 * gw_grad_sumsq writes gw_opt_scratch_doubles() partial sums; gw_adamw_ema finishes the norm, clips, updates.
 * hyper (device fp32[16], constants of the run): 0 base lr, 3 ema_decay (<0 none), 4 weight_decay, 5 max_norm (<=0 none),
 *   6 grad_scale (1/world), 7 skip_loss_threshold (<=0 off; train.py:428-436), 8 warmup_steps, 9 total_steps, 10 min_lr_scale,
 *   11 use_sched.  The LR schedule (train.py:84-91) and Adam's bias corrections are evaluated ON THE DEVICE from
 * state (device int32[4]): [0] optimisation steps applied so far, [1] batches skipped -- a skipped batch (non-finite loss or
 *   gradient norm, train.py:424-427; loss above the threshold, :428-436) leaves parameters, moments, EMA, schedule and bias
 *   correction untouched, exactly like the reference's `continue`, and no host read is needed.
 * loss: device scalar or NULL; g_has_loss != 0: the batch loss is g[n] * grad_scale instead (written by gw_bucket_reset, so it
 *   went through the same all-reduce as the gradients and every rank decides alike).
 * info (device fp32[8]) out: grad norm, clip coefficient, applied flag, lr used, loss seen. */
int gw_opt_scratch_doubles(void);
int gw_grad_sumsq(const float* g, long n, double* partial, void* stream);
int gw_adamw_ema(float* p, const float* g, float* m, float* v, float* ema, long n, const double* partial,
                 const float* hyper, const float* loss, int g_has_loss, const int* state, double beta1, double beta2,
                 float eps, float* info, void* stream);
/* g[0, n) = 0 (16-byte aligned), and g[n] = *loss when loss != NULL (the bucket's extra slot). */
int gw_bucket_reset(float* g, long n, const float* loss, void* stream);
/* *step_ctr += 1 (Philox draw counter); state[0] += 1 if info[2] (step applied) else state[1] += 1. */
int gw_train_advance(int* step_ctr, int* state, const float* info, void* stream);
/* wt[b] = (1 - alpha_bar[t_b])^power (train.py:414-417) */
int gw_loss_weight(const int64_t* t, const float* alpha_bar, float power, float* wt, int B, void* stream);

/* =====================================================================================================
 * On-device scoring of reconstructions (inference.py:11-27, 247-279, 303-314; sweep_infer.py:8-13, 225-241).
 * xhat, clean fp32 [B, L]; sigma fp32 [B] or NULL; out fp64 [B, 12]:
 *   0 corr_last, 1 mae_last (tail window t >= t_max - secs), 2 nmae_sigma (last int(fs*secs) samples), 3 overlap,
 *   4 best xcorr lag (|k| <= max_shift; <= 0 means L-1), 5 MAE / 6 NMAE_clean / 7 NMAE_sigma over [-80 ms, +40 ms] around the
 *   clean peak after alignment, 8 peak index, 9 aligned length, 10 window count, 11 tail count.
 * ===================================================================================================== */
int gw_score_batch(const float* xhat, const float* clean, const float* sigma, int B, int L, double fs, double secs,
                   int max_shift, double delta_t, double* out, void* stream);

/* =====================================================================================================
 * Shape-generic CUDA-core path (csrc/generic.cu) for the UNet1D configurations the specialised kernels do not cover:
 * base_ch not a multiple of 64 and / or kernel in {1, 3, 5, 7} (UNet1D arguments, models.py:78-88; --base_ch on the training CLI, train.py:641).
 * Any channel count; GroupNorm with `groups` = gcd(8, C) groups (models.py:160-164); stats fp32 [B, 8, 2] = (mean, rstd).
 * Same operand conventions as the specialised entry points: activations channels-last [B, L, C] in `dtype`, weights in the
 * reference layout [Cout, Cin, K], the network input fp32 [B, Cx, L] (ping-pong pair selected by *step_ptr & 1).
 *   gw_gen_conv        replaces nn.Conv1d(Cin, Cout, K, padding=K/2) of _conv_block (models.py:166-173) on
 *                      cat[nearest-upsample x2 (src0), src1] (models.py:217-222; zero rows beyond 2 L0), or on x when src0 == NULL
 *   gw_gen_gn_stats    GroupNorm statistics of the stored conv output (biased variance, eps 1e-5)
 *   gw_gen_gn_apply    GroupNorm affine + SiLU + cond 1x1 conv + FiLM (+ avg_pool1d(2, 2)) (models.py:188-193, 205-208)
 *   gw_gen_final       final Conv1d(C + 1, 1, K) on cat[h, x_t] (models.py:227-230) (+ the gw_final_step update when mode = 1)
 *   gw_gen_final_bwd, gw_gen_gn_bwd, gw_gen_wgrad, gw_gen_weight_dgrad, gw_gen_split_cat: their backward passes; parameter
 *                      gradients are ACCUMULATED (+=), dfilm rows are written.
 * ===================================================================================================== */
int gw_gen_conv(const void* src0, int C0, int L0, int up0, const void* src1, int C1, const float* x, const float* x_alt,
                const int* step_ptr, int Cx, int B, int L, const float* w, const float* bias, int Cout, int K, void* out,
                int dtype, void* stream);
int gw_gen_weight_dgrad(const float* w, int Cout, int Cin, int K, float* wt, void* stream);
int gw_gen_gn_stats(const void* raw, int B, int L, int C, int groups, int dtype, float* stats, void* stream);
int gw_gen_gn_apply(const void* raw, const float* stats, int B, int L, int C, int groups, const float* gn_w, const float* gn_b,
                    const float* cond, int Cc, const float* wc, const float* bc, const float* film, int film_off,
                    long film_b_stride, long film_step_stride, const int* step_ptr, void* out, void* pooled, int dtype,
                    void* stream);
int gw_gen_final(const void* h, int dtype, const float* net_a, const float* net_b, int B, int Cx, int L, int C, int K,
                 const float* wf, const float* bf, const gw_step_params* p, const float* coef, const int* step_ptr,
                 const float* noise, float* eps_out, float* x0_out, void* stream);
int gw_gen_final_bwd(const float* d_eps, const void* h, int dtype, const float* net, int B, int Cx, int L, int C, int K,
                     const float* wf, void* d_h, float* dWf, float* dbf, void* stream);
long gw_gen_gn_bwd_scratch_floats(int B, int C);
int gw_gen_gn_bwd(const void* raw, const float* stats, int B, int L, int C, int groups, const float* gn_w, const float* gn_b,
                  const float* cond, int Cc, const float* wc, const float* bc, const float* film, int film_off,
                  long film_b_stride, const void* d_out, const void* d_pool, int dtype, float* scratch, float* dfilm,
                  long dfilm_stride, void* d_raw, float* d_gn_w, float* d_gn_b, float* d_wc, float* d_bc, float* d_bias,
                  void* stream);
int gw_gen_wgrad(const void* src0, int C0, int L0, int up0, const void* src1, int C1, const float* x, int Cx,
                 const void* d_raw, int B, int L, int Cout, int K, int dtype, float* dW, void* stream);
int gw_gen_split_cat(const void* d_cat, int B, int L, int C0, int L0, int C1, void* d_h, void* d_skip, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif
