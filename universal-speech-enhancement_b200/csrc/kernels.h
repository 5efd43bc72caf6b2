// Host-side launchers of every CUDA kernel on the path.  All pointers are device pointers, all tensors
// are NHWC ("channels last"), all launches go to the given stream and never allocate.
// act dtype code: 0 = fp32 storage / TF32 MMA, 1 = bf16 storage / bf16 MMA.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace use {

enum ActDtype { kF32 = 0, kBF16 = 1 };
inline size_t act_size(int dt) { return dt == kBF16 ? 2 : 4; }

// ---- GroupNorm ---------------------------------------------------------------------------------
// stats: int64 fixed point [B][C][2] (sum * 2^28, sum of squares * 2^24 per channel), ACCUMULATED into (zero it first).
// Integer atomics only: bit-reproducible, independent of launch geometry and batch size.
void launch_gn_stats(int dt, const void* x, long long* stats, int B, int HW, int C, cudaStream_t st);

struct GnSrc {
  const void* x;        // act [B][Hin][Win][C]
  const long long* stats;  // fixed point [B][C][2]
  int C;
};
// out_act = FIR(silu?(groupnorm(cat[s0,s1])))  (fir: 0 none, 1 down x2, 2 up x2), written as an MMA operand
// (TF32-rounded in fp32 mode) when as_operand
// out_raw (optional, fir != 0 only) = FIR(cat[s0,s1]) un-normalised.
// aff (optional, FIR forms with a single source): the scale / shift table of launch_gn_affine; without it every tile
// recomputes its channels' scale / shift from the statistics
void launch_gn_apply(int dt, GnSrc s0, GnSrc s1, const float* gamma, const float* beta, float eps, int fir, bool do_silu,
                     bool as_operand, void* out_act, void* out_raw, int B, int Hin, int Win, cudaStream_t st,
                     const float* aff = nullptr);

// GroupNorm scale / shift table of cat[s0, s1] for the fused conv operand: aff[b][0][c] = gamma[c] * rstd,
// aff[b][1][c] = beta[c] - mean * gamma[c] * rstd (same arithmetic as launch_gn_apply)
void launch_gn_affine(GnSrc s0, GnSrc s1, const float* gamma, const float* beta, float eps, int HW, float* aff, int B,
                      cudaStream_t st);  // (PDL by pdl_enabled(B))

// ---- small / bandwidth-bound convolutions ---------------------------------------------------------
// 3x3 pad 1, C_in = 4 (fp32 NHWC input) -> N channels of act dtype.  w: [N][4][3][3] fp32, bias [N].
void launch_conv_in4(int dt, const float* x, const float* w, const float* bias, void* out, int B, int H, int W, int N,
                     cudaStream_t st);
// 3x3 pad 1, C channels (act dtype, already normalised) -> 4 fp32 channels; w: [4][C][3][3] fp32.
// If prev != nullptr adds FIR-upsample-x2(prev) where prev is fp32 [B][H/2][W/2][4].
void launch_conv_out4(int dt, const void* a, const float* w, const float* bias, const float* prev, float* out, int B,
                      int H, int W, int C, cudaStream_t st);
// out = h + bias + conv1x1_{pc->C}(pyr): Combine(method="sum").  w: [C][pc] fp32, pyr fp32 [.][pc].  In place allowed.
// stats (optional): fixed-point GroupNorm statistics [B][C][2] of `out`, accumulated (zero it first).
void launch_combine(int dt, const void* h, const float* pyr, const float* w, const float* bias, void* out, long long* stats,
                    int B, int HW, int C, int pc, cudaStream_t st);
// FIR [1,3,3,1] downsample x2 of an fp32 pc-channel map (pc = 2 or 4).
void launch_fir4_down(const float* x, float* out, int B, int Hin, int Win, int pc, cudaStream_t st);
void launch_upfirdn2d(const float* in, float* out, int major, int in_h, int in_w, int minor, const float* kernel, int kh,
                      int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                      cudaStream_t st);
// generic direct convolution (debug / verification of the tcgen05 kernel; never on the hot path)
void launch_conv_ref(int dt, const void* x, const float* w, const float* bias, int bias_bstride, const void* res,
                     float scale, void* out, int B, int H, int W, int Cin, int Cout, int ksize, cudaStream_t st);

// 3xTF32 parity mode: hi = rne_tf32(x[:, c0:c0+C]), lo = rne_tf32(x - hi), both dense [npix][C] fp32
void launch_split_tf32(const float* x, int Ct, int c0, int C, float* hi, float* lo, size_t npix, cudaStream_t st);

// ---- network input / output, SDE arithmetic -------------------------------------------------------
// xr[b][f][t][:] = 2*[Re x, Im x, Re Y, Im Y] - 1 (fp32 x4); xpad (optional): the same 4 values as act-dtype MMA
// operands zero-padded to one 128-byte channel chunk per pixel (input of the tcgen05 input convolution)
// pc = 6 (condition="both"): Y2 = the second conditioning spectrogram; xr = a [n][4] block followed by a [n][2] block
void launch_pack_input(int dt, int pc, const float2* x, const float2* Y, const float2* Y2, float* xr, void* xpad, size_t n,
                       cudaStream_t st);

// kStepDrift: no update -- x_mean receives the reverse-time drift theta (Y - x) - g^2 score pf itself (RSDE.sde()[0],
// sdes.py:122-150; the right-hand side of the probability-flow ODE when pf = 0.5)
enum StepMode { kStepReverseDiffusion = 0, kStepEulerMaruyama = 1, kStepDrift = 2 };
struct StepArgs {
  const float* pyramid;  // fp32 [B][F][T][pc]  (pc = 6: the first four channels, [B][F][T][4])
  const float* pyramid2; // pc = 6 only: channels 4, 5 as fp32 [B][F][T][2]
  int pc;                // pyramid channels: 4 (score network), 2 (discriminative network) or 6 (condition="both")
  float out_sign;        // -1: score = -net(x) (ScoreModel.forward_score); +1: the raw network output
  const float* t;        // time value of sample b at t[b * t_bstride] (divides the pyramid: scale_by_sigma) or nullptr
  int t_bstride;         // 1 = per-sample times, 0 = one batch-uniform time
  const float* ow;       // output_layer weight [2][pc] (pc <= 6)
  const float* ob;       // output_layer bias [2]
  float2* score;         // optional out: -net(x)   (ScoreModel.forward)
  // fused ReverseDiffusionPredictor step (all optional as a group; enabled when x != nullptr)
  const float2* x;
  const float2* Y;
  const float2* z;       // explicit noise or nullptr -> Philox
  float2* x_mean;
  float2* x_next;
  float theta, dt, G;    // G: reverse diffusion g(t_i) sqrt(1/N); Euler-Maruyama: g(t_i)
  float Gz;              // noise gain: G (reverse diffusion), g sqrt(1/N) (Euler-Maruyama), 0 for the probability flow
  float pf;              // 1, or 0.5 for the probability-flow ODE (sdes.py:139,166)
  int mode;              // kStepReverseDiffusion | kStepEulerMaruyama
  unsigned long long seed;
  unsigned int step;
  unsigned int clip0;    // global index of sample 0 (shard-invariant RNG streams)
  int B;
  size_t per_clip;       // F*T
};
void launch_final_step(const StepArgs& a, cudaStream_t st);
// One corrector update (LangevinCorrector / AnnealedLangevinDynamics, sampling/correctors.py:37-98) given grad = score(x):
// x_mean = x + step grad ; x = x_mean + z sqrt(2 step).  langevin: step = 2 (snr mean_b||z_b|| / mean_b||grad_b||)^2 is
// reduced on the device (deterministic two-pass sum); otherwise `step` is the host-computed 2 (snr std(t))^2 of ALD.
struct CorrectorArgs {
  const float2* x;
  const float2* grad;
  const float2* z;        // explicit noise or nullptr -> Philox stream `draw`
  float2* x_mean;
  float2* x_next;
  int langevin;
  float snr, step;
  void* scratch;          // corrector_scratch_bytes(B) device bytes (Langevin only)
  unsigned long long seed;
  unsigned int draw, clip0;
  int B;
  size_t per_clip;
};
// forward half of train_step: coef = device fp32 [2][B] = (exp(-theta t_b), std(t_b)); z explicit or Philox stream 0xfffffffe
void launch_perturb(const float2* X0, const float2* Y, const float2* z, const float* coef, float2* xt, unsigned long long seed,
                    unsigned int clip0, int B, size_t per_clip, cudaStream_t st);
// loss[0] = mean_b 0.5 sum |score std + z|^2 (mae: |.|), loss[1 + b] = the per-clip terms; scratch: corrector_scratch_bytes(B)
void launch_dsm_loss(const float2* score, const float2* z, const float* coef, void* scratch, float* loss, int mae,
                     unsigned long long seed, unsigned int clip0, int B, size_t per_clip, cudaStream_t st);
size_t corrector_scratch_bytes(int B);
void launch_corrector_step(const CorrectorArgs& c, cudaStream_t st);
// x0 = Y + z * std   (z explicit or Philox, step = 0xffffffff stream)
void launch_prior(const float2* Y, const float2* z, float2* x0, float std, unsigned long long seed, unsigned int clip0,
                  int B, size_t per_clip, cudaStream_t st);
// fill z with the Philox complex normal stream of (seed, step, clip0 + b) -- test hook for the RNG
void launch_philox_fill(float2* z, unsigned long long seed, unsigned int step, unsigned int clip0, int B,
                        size_t per_clip, cudaStream_t st);

// Programmatic dependent launch (conv_tc / gn_affine chains, common.cuh): the successor's set-up overlaps its predecessor's
// tail.  Measured (bf16, CUDA graphs, round 2, runs repeat to 0.01 ms): batch 1 148.8 -> 145.4 ms per clip, batch 4
// 388.3 -> 393.3 ms per step: a latency optimisation, so it is on for launches of at most kPdlMaxBatch clips.
// USE_B200_PDL=0 / 1 forces it off / on, USE_B200_PDL_MAXB=<n> moves the threshold.
bool pdl_enabled(int B);

// ---- time embedding ---------------------------------------------------------------------------------
// gfp (Fourier features, sample b at gfp + b * gfp_bstride) -> silu(Linear(silu(Linear(gfp)))) [B][4*nf]
void launch_temb_mlp(const float* gfp, int gfp_bstride, const float* w1, const float* b1, const float* w2, const float* b2, float* out,
                     int B, int nf, cudaStream_t st);
// out[b][n] = base[n] + sum_k W[n][k] * temb[b][k]  for all rows of all ResBlocks at once
// (temb row of sample b at temb + b * temb_bstride; stride 0 = one shared row)
void launch_dense_all(const float* temb, int temb_bstride, const float* W, const float* base, float* out, int B, int rows,
                      int K, cudaStream_t st);

// ---- attention block (bottleneck only) ----------------------------------------------------------------
// q, k, v = NIN_0/1/2(in) in one launch: in is [M][C] in the activation dtype, W [C][C] ([in][out]), outputs fp32 [M][C]
void launch_nin_qkv(int dt, const void* in, const float* W0, const float* b0, float* q, const float* W1, const float* b1,
                    float* k, const float* W2, const float* b2, float* v, int M, int C, cudaStream_t st);
// softmax(q k^T / sqrt(C)) v per sample over its P positions, fp32
void launch_attn_core(const float* q, const float* k, const float* v, float* out, int B, int P, int C, cudaStream_t st);
// out_act = (x + NIN_3(att)) * scale
void launch_nin_proj(int dt, const float* att, const float* W, const float* b, const void* x, float scale, void* out, int M,
                     int C, cudaStream_t st);

// ---- STFT front / back end ------------------------------------------------------------------------------
// y [B][L] -> Y complex [B][F][Tp] with spectral compression (|S|^e e^{j angle} * factor), zero for frames >= T
void launch_stft(const float* y, float2* Y, const float* window, const float2* twiddle, int B, int L, int n_fft, int hop,
                 int T, int Tp, float factor, float exponent, cudaStream_t st);
// X complex [B][F][Tp] -> y [B][L]; frames scratch fp32 [B][Tp][n_fft]
void launch_istft(const float2* X, float* frames, float* y, const float* window, const float2* twiddle,
                  const float* inv_env, int B, int L, int n_fft, int hop, int Tp, float factor, float exponent,
                  cudaStream_t st);

// ---- predict-side audio preparation (resample.cu) ---------------------------------------------------------------------
// scipy.signal.resample / librosa res_type="fft" semantics: x [B][n_in] -> y[b * y_stride + j], j < n_out
size_t resample_workspace_bytes(int B, int n_in, int n_out);
void launch_resample_fft(const float* x, float* y, int B, int n_in, int n_out, int y_stride, void* work, cudaStream_t st);
// y [B][stride]: clip b keeps its first lengths[b] samples scaled to peak `target` (<= 0: unscaled), the rest is zeroed
void launch_peak_normalize_pad(float* y, const int* lengths, int B, int stride, float target, unsigned int* peaks,
                               cudaStream_t st);

// ---- tcgen05 convolution ----------------------------------------------------------------------------------
struct TcSegDesc {
  const void* act;   // act tensor [B][H][W][C_tensor]
  int C_tensor;      // channels of the tensor (row pitch)
  int c0, C;         // channel window used by this segment
  const void* w;     // packed weights [taps][N][Cw_total] (act dtype)
  int Cw_total;      // total input channels of the weight tensor
  int wc0;           // first weight channel used by this segment
  int taps;          // 9 or 1
  // fused GroupNorm + SiLU operand (3x3 only): `act` is the RAW tensor and the kernel normalises it on the way into
  // shared memory with the per-sample table aff = fp32 [B][2][aff_C] (scale row, shift row; launch_gn_affine);
  // channel c0 of `act` uses column aff_c0 of the table.  nullptr: `act` is consumed as is.
  const float* aff;
  int aff_C, aff_c0;
};
struct TcConvDesc {
  TcSegDesc seg[3];
  int nseg;
  // inline GroupNorm: when gn_st0 != nullptr the fused segments' scale / shift table (GroupNorm over cat[s0, s1] with the
  // fixed-point statistics st0 [B][C0][2], st1 [B][C1][2]) is computed INSIDE the convolution kernel; the segments then only
  // use aff_c0 (their first channel inside the concatenation) and must set aff to any non-null value
  const long long* gn_st0;
  const long long* gn_st1;
  int gn_C0, gn_C1, gn_HW;
  const float* gn_gamma;
  const float* gn_beta;
  float gn_eps;
  int B, H, W, N;
  void* out;
  const float* bias;
  int bias_bstride;
  const void* res;
  float scale;
  long long* stats_acc;  // optional fixed-point GroupNorm statistics of `out`, [B][N][2], zero on entry
  // latency kernels: the low-resolution levels (<= 3 tiles per clip) run the split-K cluster form (conv_tc_ks.cuh), whose
  // partial-sum association differs from the single-accumulator form in the last bits
  int latency;
};
struct TcConvPlan;  // opaque: tensor maps + launch geometry
// Build (host) the launch plan; returns nullptr and fills err on failure.
TcConvPlan* tc_conv_plan_create(int dt, const TcConvDesc& d, int num_sms, char* err, int errlen);
void tc_conv_plan_destroy(TcConvPlan* p);
void tc_conv_launch(const TcConvPlan* p, cudaStream_t st);
bool tc_conv_supported(int dt, int N);

// ---- pyramid head: 3x3 pad 1, C -> pc (4 or 2) fp32 channels (+ FIR-upsampled previous pyramid), nine taps folded into
// the MMA's N dimension (head_tc.cuh).  w_packed: act dtype [48][C], row tap * pc + co (use_pack_head_weight).
struct HeadPlan;
bool head_tc_supported(int dt, int C, int pc);
// aff != nullptr: `act` is the RAW tensor and aff its GroupNorm scale / shift table fp32 [B][2][C] (launch_gn_affine):
// normalise + SiLU + operand rounding happen inside the kernel (fused operand, bit-identical to gn_apply + plain head)
HeadPlan* head_tc_plan_create(int dt, const void* act, const void* w_packed, const float* bias, const float* prev4,
                              float* out4, int B, int H, int W, int C, int pc, int num_sms, char* err, int errlen,
                              const float* aff = nullptr);
void head_tc_plan_destroy(HeadPlan* p);
void head_tc_launch(const HeadPlan* p, cudaStream_t st);
int tc_conv_tiles_per_image(int dt, int N, int H, int W);

}  // namespace use
