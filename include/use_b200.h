/* use_b200 -- C ABI of the B200-native SGMSE sampling path (libuse_b200.so).
 *
 * Plain pointers and sizes only; no torch types.  Every function returns 0 on success and a non-zero
 * code on failure, with a message retrievable through use_last_error() (the reference surfaces native
 * failures as Python RuntimeError through TORCH_CHECK, op/upfirdn2d.cpp:8-10; the Python host layer maps
 * non-zero codes to RuntimeError the same way).  Inputs are borrowed and never mutated; outputs and all
 * scratch memory are caller-provided (the reference allocates outputs with the input's options inside
 * the op, upfirdn2d_kernel.cu:242-243 -- here the caller's allocator, torch's, stays the only one).
 * All work is enqueued on the cudaStream_t passed as `stream` (the reference launches on the current
 * stream of the current device, upfirdn2d_kernel.cu:213-215) and the functions are re-entrant per
 * stream for distinct engines.
 *
 * What each entry point replaces in /root/reference (paths relative to src/models/components/sgmse/):
 *   use_upfirdn2d_f32        backbones/ncsnpp_utils/op/upfirdn2d.cpp:12-23 (pybind `upfirdn2d`), the
 *                            reference's only native FFI on the path; kernels op/upfirdn2d_kernel.cu:107-207
 *   use_engine_*             construction of NCSNpp (backbones/ncsnpp.py:42-316) + load of its state_dict
 *   use_score_forward        ScoreModel.forward / forward_score (model_wrapper.py:135-145) -> NCSNpp.forward
 *                            (backbones/ncsnpp.py:324-501)
 *   use_pc_sample            sampling.get_pc_sampler().pc_sampler (sampling/__init__.py:59-71) with
 *                            ReverseDiffusionPredictor (sampling/predictors.py:61-68), NoneCorrector
 *                            (sampling/correctors.py:101-111), RSDE.discretize (sdes.py:159-173)
 *   use_pc_sample_ex         the same loop with EulerMaruyamaPredictor (predictors.py:40-53), LangevinCorrector /
 *                            AnnealedLangevinDynamics (sampling/correctors.py:37-98), probability flow, denoise=False
 *   use_reverse_drift        RSDE.sde()[0] (sdes.py:122-150): the drift function of get_ode_sampler
 *                            (sampling/__init__.py:76-159)
 *   use_train_forward        forward half of ScoreModel.train_step (model_wrapper.py:147-208): marginal_prob perturbation,
 *                            one score evaluation with per-sample times, denoising-score-matching loss (:124-133)
 *   use_stft / use_istft     ScoreModel.stft + spec_fwd + pad_spec / spec_back + istft
 *                            (model_wrapper.py:92-122, util/other.py:128-135)
 *   use_op_*                 single kernels, exported for the parity tests
 */
#ifndef USE_B200_H_
#define USE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define USE_B200_ABI_VERSION 5

#define USE_DTYPE_F32 0  /* fp32 storage, TF32 tensor-core math (PyTorch's own GPU default for conv) */
#define USE_DTYPE_BF16 1 /* bf16 storage + bf16 tensor-core math, fp32 accumulate / statistics / SDE state */
#define USE_DTYPE_F32X3 2 /* parity mode: fp32 storage, every tensor-core convolution evaluated as the 3xTF32 split
                           * x_hi w_hi + x_hi w_lo + x_lo w_hi (fp32-level accuracy at a third of the TF32 rate); engine
                           * only (use_config.act_dtype), the single-kernel exports take USE_DTYPE_F32 / _BF16 */

typedef struct use_engine use_engine;

typedef struct use_config {
  int nf;               /* base width (NCSNppLarge: 128) */
  int num_levels;       /* len(ch_mult) (7) */
  int ch_mult[8];       /* (1,1,2,2,2,2,2) */
  int num_res_blocks;   /* 2 */
  int input_channels;   /* 4 = [Re x, Im x, Re Y, Im Y] (score network); 2 = [Re x, Im x] (discriminative network);
                         * 6 = [x, Y, Y2] (condition="both", model_wrapper.py:43-46: noisy + GAN-denoised conditioning) */
  int act_dtype;        /* USE_DTYPE_* */
  int n_fft, hop;       /* 1022, 160 */
  float spec_factor;    /* 0.15 */
  float spec_abs_exponent; /* 0.5 */
  float theta;          /* OUVE stiffness 1.5 */
  int conditional;      /* 1: time embedding -> per-ResBlock bias (score network); 0: NCSNpp(discriminative=True) */
  int scale_by_sigma;   /* 1: the output pyramid is divided by the time value (ncsnpp.py:492-494) */
} use_config;

int use_abi_version(void);
const char* use_last_error(void);

/* ---- engine lifetime and weights ------------------------------------------------------------- */
use_engine* use_engine_create(const use_config* cfg);
void use_engine_destroy(use_engine* e);
/* Register one state_dict tensor (host fp32, key relative to score_net, e.g. "all_modules.4.Conv_0.weight"). */
int use_engine_set_weight(use_engine* e, const char* name, const float* host, const int64_t* shape, int ndim);
/* Validate that every tensor of the architecture is present and pack them (host side); returns the number of
 * device bytes the packed blob needs through *bytes. */
int use_engine_pack(use_engine* e, size_t* bytes);
/* Copy the packed blob into caller-allocated device memory (kept by the engine until destroy). */
int use_engine_upload(use_engine* e, void* dev_weights, size_t bytes, void* stream);
/* Workspace (device bytes) a forward / sample over B spectrograms of F x T (T % 2^(levels-1) == 0) needs. */
int use_engine_workspace_bytes(use_engine* e, int B, int F, int T, size_t* bytes);
/* Options: "overlap_groups" = 1 | 2 (default 2): use_pc_sample splits an even batch >= 4 into two halves that run
 * on their own streams so the HBM-bound kernels of one half overlap the tensor-bound convolutions of the other.
 * A/B switches whose two settings give bit-identical results: "fuse_gn" (GroupNorm + SiLU inside the convolution's
 * operand path), "fuse_head" (the same inside the pyramid heads), "use_graphs" (CUDA-graph replay of an evaluation),
 * "inline_gn" (GroupNorm scale / shift tables computed inside the consumer kernels instead of one gn_affine_kernel
 * launch per GroupNorm).
 * "ksplit" = 0 | 1 | 2 (default 2 = auto): LATENCY mode.  The convolutions of the low-resolution levels (<= 32 x 40
 * pixels per clip) run as split-K clusters of 2 / 4 CTAs with a DSMEM reduction in rank order: batch 1 -6 % (bf16) /
 * -9 % (TF32), but 2 % slower from ~16 clips on.  auto = calls of at most two clips.  The mode is a property of the whole
 * launch program: inside a mode per-clip results are bit-identical for every batch size; the two modes differ in the
 * last bits (same tolerance to the reference).  A caller that splits one job over several calls (micro-batches, ranks)
 * sets 0 / 1 explicitly from the size of the whole job (the Python layer does: ScoreModel.sample(job_clips=...)). */
int use_engine_set_option(use_engine* e, const char* key, int value);

/* Instrumentation: kernels launched so far by this engine; per-op-class CUDA-event timing of network evaluations
 * (events on the launch stream, one pair per op; enable only outside timed regions). */
long long use_engine_launch_count(use_engine* e);
int use_engine_set_profiling(use_engine* e, int on);
int use_engine_get_profile(use_engine* e, char* json, size_t cap);
int use_engine_get_profile_ops(use_engine* e, char* csv, size_t cap);

/* ---- the hot path ------------------------------------------------------------------------------ */
/* score = -net(cat[x, Y], t).  x, Y, score: device complex64 [B][F][T] (interleaved re, im).
 * t_host [B] and gfp_host [B][2*nf] = [sin(2 pi W log t), cos(..)] are HOST arrays (the Fourier features are
 * evaluated by the host layer with torch so the embedding is bit-identical to the reference's). */
int use_score_forward(use_engine* e, int B, int F, int T, const void* x, const void* Y, const float* t_host,
                      const float* gfp_host, void* score, void* workspace, size_t workspace_bytes, void* stream);

/* The same for the 6-channel network (condition="both"): score = -net(cat[x, Y, Y2], t). */
int use_score_forward2(use_engine* e, int B, int F, int T, const void* x, const void* Y, const void* Y2, const float* t_host,
                       const float* gfp_host, void* score, void* workspace, size_t workspace_bytes, void* stream);

/* Reverse-time drift of the SDE / probability-flow ODE at a batch-uniform time t (RSDE.sde(x, t, y)[0], sdes.py:122-150):
 *   drift = theta (sde_y - x) - g^2 score(x, t) c,  c = 1/2 with probability_flow (the ODE's right-hand side,
 *   sampling/__init__.py:112-114), 1 otherwise.  cond / cond2: the network's conditioning when it is not sde_y / the
 * 6-channel network's second conditioning (NULL otherwise).  g = g(t) from the host (OUVESDE.sde).  One network
 * evaluation + one fused kernel; the ODE solver itself (scipy RK45 in the reference) stays on the host. */
int use_reverse_drift(use_engine* e, int B, int F, int T, const void* x, const void* sde_y, const void* cond,
                      const void* cond2, const float* t_host, const float* gfp_host, float g, int probability_flow, void* drift,
                      void* workspace, size_t workspace_bytes, void* stream);

/* out = +net(...): NCSNpp.forward itself (ncsnpp.py:324-501).  For the discriminative generator of the LSGAN stage
 * (GAN/generator/ncsnpp/model_wrapper.py:54,114-121: input_channels = 2, conditional = 0, scale_by_sigma = 0) pass
 * Y = t_host = gfp_host = NULL. */
int use_net_forward(use_engine* e, int B, int F, int T, const void* x, const void* Y, const float* t_host,
                    const float* gfp_host, void* out, void* workspace, size_t workspace_bytes, void* stream);

/* N reverse-diffusion predictor steps.  Y: device complex64 [B][F][T]; x_mean (out) same shape: the
 * noise-free mean of the last step (denoise=True).  x_state: device scratch of the same shape.
 * t_host[N], G_host[N] (= g(t_i) sqrt(1/N)), gfp_host[N][2*nf]: the float32 step schedule, host arrays.
 * noise: device complex64 [N+1][B][F][T] explicit draws (prior, then one per step) or NULL for the in-kernel
 * Philox stream keyed by (seed, step, clip0 + b).  */
int use_pc_sample(use_engine* e, int B, int F, int T, const void* Y, void* x_state, void* x_mean, int N,
                  const float* t_host, const float* G_host, const float* gfp_host, float prior_std, const void* noise,
                  uint64_t seed, uint32_t clip0, void* workspace, size_t workspace_bytes, void* stream);

/* The general predictor-corrector sampler (sampling/__init__.py:23-73): use_pc_sample with the other registered
 * predictors / correctors fused into the same one-call loop.
 *   predictor  reverse_diffusion (predictors.py:56-68) | euler_maruyama (predictors.py:40-53 over RSDE.rsde_parts,
 *              sdes.py:128-150) | none (predictors.py:71-79)
 *   corrector  none (correctors.py:101-111) | langevin (correctors.py:37-64; its batch-mean norms are a deterministic
 *              two-pass device reduction, the batch is never split into stream groups) | ald (correctors.py:67-98)
 * Normal draws are consumed in the reference's order: the prior, then per outer step `corrector_steps` corrector draws
 * and one predictor draw; explicit `noise` is complex64 [1 + N * (corrector_steps_eff + predictor_eff)][B][F][T]
 * (corrector_steps_eff = 0 for corrector none, predictor_eff = 0 for predictor none), the Philox stream of draw d is
 * keyed (seed, d, clip0 + b).  opts == NULL: reverse_diffusion + none, denoise = 1 (= use_pc_sample). */
#define USE_PRED_REVERSE_DIFFUSION 0
#define USE_PRED_EULER_MARUYAMA 1
#define USE_PRED_NONE 2
#define USE_CORR_NONE 0
#define USE_CORR_LANGEVIN 1
#define USE_CORR_ALD 2
typedef struct use_sampler_opts {
  int predictor;              /* USE_PRED_* */
  int corrector;              /* USE_CORR_* */
  int corrector_steps;        /* n_steps of the corrector */
  float snr;                  /* Langevin target SNR */
  int probability_flow;       /* 1: score term halved, no predictor noise (sdes.py:139-143,166-170) */
  int denoise;                /* 1: return the noise-free mean of the last step (the state itself when the predictor is
                               * "none": NonePredictor returns (x, x)); 0: the state; 2: the mean of the last update of
                               * any kind (a corrector-only step: Corrector.update_fn's own x_mean) */
  const float* g_host;        /* [N] diffusion g(t_i) (host), required by euler_maruyama */
  const float* ald_step_host; /* [N] 2 (snr std(t_i))^2 (host), required by ald */
  void* trace;                /* optional device complex64 [N][B][F][T]: xt_mean after every outer step (parity tests) */
  const void* x_init;         /* optional device complex64 [B][F][T]: start from this state instead of the prior draw
                               * (single update_fn steps of the registry classes: N = 1 tables + dt_steps = sde.N) */
  int dt_steps;               /* > 0: dt = 1 / dt_steps instead of 1 / N */
  const void* cond;           /* optional device complex64 [B][F][T]: the network's conditioning spectrogram when it is not
                               * the SDE's y (condition="denoised" with sde_input="noisy" and vice versa,
                               * model_wrapper.py:281-299); NULL: the network is conditioned on Y */
  const void* cond2;          /* 6-channel network (condition="both"): the second conditioning spectrogram (the
                               * GAN-denoised one, model_wrapper.py:287-288); required there, ignored otherwise */
} use_sampler_opts;
int use_pc_sample_ex(use_engine* e, int B, int F, int T, const void* Y, void* x_state, void* x_mean, int N,
                     const float* t_host, const float* G_host, const float* gfp_host, float prior_std, const void* noise,
                     uint64_t seed, uint32_t clip0, const use_sampler_opts* opts, void* workspace, size_t workspace_bytes,
                     void* stream);

/* Forward half of ScoreModel.train_step (model_wrapper.py:147-208; the backward pass / optimizer are out of scope):
 *   x_t  = exp(-theta t_b) X0 + (1 - exp(-theta t_b)) Y + std(t_b) z        (OUVESDE.marginal_prob, sdes.py:225-247)
 *   loss = mean_b 0.5 sum |score(x_t, t_b) std(t_b) + z|^2  (loss_type 0, "mse")  or  ... sum |.|  (1, "mae")  (:124-133)
 * X0 (clean), Y (noisy): device complex64 [B][F][T] compressed spectrograms; t_host [B] per-sample times, gfp_host
 * [B][2*nf] their Fourier features, coef_host [2][B] = (exp(-theta t_b), std(t_b)) -- host float32 arrays from the host
 * layer's torch expressions; noise: device complex64 [B][F][T] explicit z or NULL (Philox, keyed by seed and clip0 + b).
 * x_t (out): device complex64 [B][F][T]; loss (out): device float [1 + B] = (batch mean, per-clip terms).  The sums are
 * deterministic two-pass reductions (no floating-point atomics). */
int use_train_forward(use_engine* e, int B, int F, int T, const void* X0, const void* Y, const float* t_host,
                      const float* gfp_host, const float* coef_host, const void* noise, uint64_t seed, uint32_t clip0,
                      int loss_type, void* x_t, float* loss, void* workspace, size_t workspace_bytes, void* stream);

/* y: device float [B][L] -> Y: device complex64 [B][n_fft/2+1][Tp], frames >= 1 + L/hop zero (pad_spec).
 * window [n_fft] and twiddle [n_fft] (cos, sin of 2 pi i / n_fft, interleaved) are device arrays. */
int use_stft(use_engine* e, int B, int L, int Tp, const float* y, void* Y, const float* window, const float* twiddle,
             void* stream);
/* X: device complex64 [B][F][Tp] -> y device float [B][L].  frames_scratch: device float [B][Tp][n_fft].
 * envelope: device float [n_fft + hop (Tp-1)] = overlap-added squared window. */
int use_istft(use_engine* e, int B, int L, int Tp, const void* X, float* y, float* frames_scratch, const float* window,
              const float* twiddle, const float* envelope, void* stream);

/* ---- predict-side audio preparation ------------------------------------------------ */
/* librosa.resample(y, orig_sr, target_sr, res_type="fft") of LoadWavDataset.__getitem__ (src/data/components/
 * loadwav_dataset.py:94-98) = scipy.signal.resample(y, n_out) with n_out = ceil(n_in * target_sr / orig_sr): x device float
 * [B][n_in] -> y[b * y_stride + j], j < n_out.  Any lengths (Bluestein chirp-z + radix-2 Stockham passes); `work`: device
 * scratch of use_resample_workspace_bytes. */
int use_resample_workspace_bytes(int B, int n_in, int n_out, size_t* bytes);
int use_resample_fft_f32(const float* x, int B, int n_in, float* y, int n_out, int y_stride, void* work, size_t work_bytes,
                         void* stream);
/* y / max|y| * target_peak per clip over its first lengths[b] samples (loadwav_dataset.py:99-100; target_peak <= 0: no
 * scaling; an all-zero clip stays zero) and zero padding up to `stride` (collate.pad_to_longest_monaural_inference,
 * collate.py:42-73), in place.  lengths_dev: device int [B]; peaks_scratch: device, 4 * B bytes (holds max|y| as float). */
int use_peak_normalize_pad_f32(float* y, const int* lengths_dev, int B, int stride, float target_peak, void* peaks_scratch,
                               void* stream);

/* ---- the reference's native op -------------------------------------------------------------------- */
/* upfirdn2d over [major][in_h][in_w][minor] fp32 (op/upfirdn2d.cpp:12-23 argument order). */
int use_upfirdn2d_f32(const float* in, float* out, int major, int in_h, int in_w, int minor, const float* kernel,
                      int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                      int pad_y1, void* stream);

/* ---- single kernels (parity tests) ----------------------------------------------------------------- */
/* GroupNorm statistics: int64 FIXED POINT [B][C][2] = (sum * 2^28, sum of squares * 2^24) per channel, ACCUMULATED with
 * integer atomics into `stats` (zero it first): bit-reproducible, independent of launch geometry and batch size. */
#define USE_STAT_SUM_SCALE 268435456.0
#define USE_STAT_SQ_SCALE 16777216.0
int use_op_gn_stats(int dtype, const void* x, long long* stats, int B, int HW, int C, void* stream);
int use_op_gn_apply(int dtype, const void* x0, const long long* stats0, int C0, const void* x1, const long long* stats1, int C1,
                    const float* gamma, const float* beta, float eps, int fir, int do_silu, int as_operand, void* out_act,
                    void* out_raw, int B, int Hin, int Win, void* stream);
/* Same with the scale / shift table of use_op_gn_affine (fp32 [B][2][C0]) for the resampling forms (fir != 0, C1 == 0):
 * the tiles read the table instead of re-deriving their channels' scale / shift from the statistics. */
int use_op_gn_apply_aff(int dtype, const void* x0, const long long* stats0, int C0, const void* x1, const long long* stats1,
                        int C1, const float* gamma, const float* beta, float eps, int fir, int do_silu, int as_operand,
                        void* out_act, void* out_raw, int B, int Hin, int Win, const float* aff, void* stream);
/* tcgen05 implicit-GEMM convolution; up to 3 segments summed into one accumulator.
 * seg_act[i]: act tensor [B][H][W][seg_ctensor[i]], channel window [seg_c0, seg_c0+seg_c);
 * seg_w[i]: packed weights [taps][N][seg_cw[i]] in act dtype, window starting at seg_wc0[i].
 * stats (optional): fixed-point [B][N][2] GroupNorm statistics of `out` accumulated by the epilogue (zero it first). */
int use_op_conv_tc(int dtype, int nseg, const void* const* seg_act, const int* seg_ctensor, const int* seg_c0,
                   const int* seg_c, const void* const* seg_w, const int* seg_cw, const int* seg_wc0, const int* seg_taps,
                   int B, int H, int W, int N, const float* bias, int bias_bstride, const void* res, float scale, void* out,
                   long long* stats, void* stream);
/* Same convolution with FUSED GroupNorm + SiLU operands (layerspp.py:283-285,304-306 without materialising the
 * normalised tensor): a 3x3 segment whose seg_aff[i] != NULL reads the RAW tensor seg_act[i] and applies
 * silu(x * scale + shift) + operand rounding on the way into shared memory; seg_aff[i] = fp32 [B][2][seg_aff_c[i]]
 * (scale row, shift row) from use_op_gn_affine, channel seg_c0[i] of the tensor <-> column seg_aff_c0[i]. */
int use_op_gn_affine(const long long* stats0, int C0, const long long* stats1, int C1, const float* gamma,
                     const float* beta, float eps, int HW, float* aff, int B, void* stream);
int use_op_conv_tc_gn(int dtype, int nseg, const void* const* seg_act, const int* seg_ctensor, const int* seg_c0,
                      const int* seg_c, const void* const* seg_w, const int* seg_cw, const int* seg_wc0, const int* seg_taps,
                      const float* const* seg_aff, const int* seg_aff_c, const int* seg_aff_c0, int B, int H, int W, int N,
                      const float* bias, int bias_bstride, const void* res, float scale, void* out, long long* stats,
                      void* stream);
/* Test hook: the use_op_conv_tc* entry points plan their launches like a latency-mode program (split-K clusters for images
 * of at most 10 tiles, see use_engine_set_option "ksplit"). */
int use_op_set_latency(int on);
int use_op_conv_ref(int dtype, const void* x, const float* w, const float* bias, int bias_bstride, const void* res,
                    float scale, void* out, int B, int H, int W, int Cin, int Cout, int ksize, void* stream);
int use_op_conv_in4(int dtype, const float* x, const float* w, const float* bias, void* out, int B, int H, int W, int N,
                    void* stream);
int use_op_conv_out4(int dtype, const void* a, const float* w, const float* bias, const float* prev, float* out, int B,
                     int H, int W, int C, void* stream);
/* Pyramid head (ncsnpp.py:440-461): 3x3 pad 1, C -> pc (4 or 2) fp32 channels of the activated tensor a (+ FIR-upsample
 * x2 of prev, fp32 [B][H/2][W/2][pc]) on the tensor cores with the nine taps folded into the MMA's N dimension.
 * w_oihw_host: fp32 [pc][C][3][3] on the HOST; w_packed_dev: device scratch of 48 * C * sizeof(act) bytes. */
int use_op_head_tc(int dtype, const void* a, const float* w_oihw_host, const float* bias, const float* prev, float* out,
                   int B, int H, int W, int C, int pc, void* w_packed_dev, void* stream);
/* The same head over the RAW tensor x with GroupNorm + SiLU fused into its operand path (layerspp.py / ncsnpp.py:440-446:
 * `act(GroupNorm(h))` in front of the pyramid conv): aff = fp32 [B][2][C] scale row / shift row from use_op_gn_affine. */
int use_op_head_tc_gn(int dtype, const void* x, const float* aff, const float* w_oihw_host, const float* bias,
                      const float* prev, float* out, int B, int H, int W, int C, int pc, void* w_packed_dev, void* stream);
int use_op_combine(int dtype, const void* h, const float* pyr, const float* w, const float* bias, void* out, int B, int HW,
                   int C, int pc, void* stream);
/* Combine that also accumulates the fixed-point GroupNorm statistics [B][C][2] of its output (zero `stats` first). */
int use_op_combine_stats(int dtype, const void* h, const float* pyr, const float* w, const float* bias, void* out,
                         long long* stats, int B, int HW, int C, int pc, void* stream);
int use_op_fir4_down(const float* x, float* out, int B, int Hin, int Win, int pc, void* stream);
int use_op_philox(void* z, uint64_t seed, uint32_t step, uint32_t clip0, int B, size_t per_clip, void* stream);
/* pack fp32 OIHW conv weights into the tcgen05 layout [taps][O][I] of the act dtype (host -> host). */
int use_pack_conv_weight(int dtype, const float* w_oihw, int O, int I, int ksize, void* out);
/* pack fp32 [pc][C][3][3] pyramid-head weights into the head kernel's layout: act dtype [48][C], row tap * pc + c_out
 * (rows >= 9 * pc are zero): the nine taps folded into the MMA's N dimension (host -> host). */
int use_pack_head_weight(int dtype, const float* w_oihw, int pc, int C, void* out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* USE_B200_H_ */
