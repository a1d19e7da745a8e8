// Bandwidth-bound kernels around the GEMMs: InstanceNorm apply / backward,
// reflection-pad writes, activation backward, image re-layout, fused losses,
// Keras-Adam. All activations are bf16 NHWC with C % 8 == 0 (16-byte vectors).
#pragma once
#include "common.cuh"

namespace b2 {

// Strided NHWC view (channel stride 1). `p` points at logical pixel (0, 0) of
// image 0, so halo pixels have negative h / w.
struct View {
  void* p;
  long long sn, sh, sw;
};

struct InApplyParams {
  View x;               // raw conv output (bf16)
  const float* stats;   // [N][C][2] sum / sumsq, or nullptr (no normalisation)
  const float* gamma;   // [C]
  const float* beta;    // [C]
  float inv_count;      // 1 / (H*W)
  float eps;
  int act;
  View skip;            // residual input (bf16) or p == nullptr
  View dst;             // bf16 destination
  int halo;             // reflect-pad width written around dst (0 = none)
  int N, H, W, C;
  // optional FP8 (E4M3) copy of the output for an FP8 consumer conv: bytes = sat(v * *scale8);
  // |v| maxima are accumulated into *amax8 (delayed scaling). Same logical layout / halo as dst.
  View dst8;            // uint8 view (strides in bytes), p == nullptr if unused
  const float* scale8;
  float* amax8;
  // MXFP8 twin instead (sf8.p != nullptr, C == 256): dst8 gets E4M3 bytes scaled per 32-channel block and
  // sf8 (uint8 view, 16-byte records per pixel, same halo) the E8M0 block scales; scale8 / amax8 unused.
  View sf8;
};
void launch_in_apply(const InApplyParams& p, cudaStream_t s);
// The InstanceNorm kernels stream their inputs through a shared-memory ring filled by the bulk-copy
// engine (default); `set_ew_stream(false)` selects the direct (register-landing) kernels, which also
// serve the FP8-twin output and views whose pixel stride is not C. For A/B tests and benchmarks.
void set_ew_stream(bool on);
bool ew_stream();

struct InBwdReduceParams {
  View g1;              // bf16 gradient source 1 (may be a padded-domain dgrad output)
  int fold1;            // reflect-pad width to fold back from g1's halo (0 = plain)
  View g2;              // optional second source (plain), p == nullptr if unused
  View x;               // raw conv output of this layer (bf16)
  const float* stats;   // forward statistics [N][C][2]; nullptr => layer has no IN (dy = g*act')
  const float* gamma;
  const float* beta;
  float inv_count, eps;
  int act;              // activation that followed the norm
  View dy;              // bf16 out: g * act'(y)   (== g when act is none)
  float* bstats;        // [N][C][2]: sum dy, sum dy*xhat (atomics), may be nullptr when !stats
  int N, H, W, C;
};
void launch_in_bwd_reduce(const InBwdReduceParams& p, cudaStream_t s);

struct InBwdApplyParams {
  View dy;              // bf16
  View x;               // raw conv output
  const float* stats;
  const float* bstats;
  const float* gamma;
  float inv_count, eps;
  View dx;              // bf16 out (interior of a zero-ringed buffer or plain)
  float* dgamma;        // fp32 grads, accumulated (+=)
  float* dbeta;
  int N, H, W, C;
  // optional MXFP8 twin of dx for an MXFP8 dgrad (C == 256, streamed kernel): E4M3 bytes + 16-byte scale
  // records per pixel, same geometry as dx (the zero ring of the flat dgrad domain stays zero)
  View dx8, sf8;
};
void launch_in_bwd_apply(const InBwdApplyParams& p, cudaStream_t s);

// dy = (g1 + g2) * act'(out) for layers without IN (PatchGAN conv0: leaky; generator
// final conv: tanh), plus the bias gradient. `out` is the stored post-activation tensor.
struct ActBwdParams {
  View g1;
  int fold1;
  View g2;
  View out;             // post-activation output (bf16)
  int act;
  View dy;              // bf16 out
  float* dbias;         // fp32 [C_valid] accumulated, may be nullptr
  int c_valid;          // channels that carry data (3 of 8 for images)
  int N, H, W, C;
};
void launch_act_bwd(const ActBwdParams& p, cudaStream_t s);

// Image -> 8-channel bf16 layouts: reflect-padded (generator stem input) and
// zero-padded (PatchGAN conv0 input). Source is fp32 NHWC3 or bf16 NHWC8.
struct ImgPrepParams {
  const void* src;
  int src_is_f32;       // 1: fp32 [N,H,W,3]; 0: bf16 view with 8 channels (3 valid)
  View src_view;        // used when !src_is_f32
  View dst_reflect;     // p == nullptr to skip
  int reflect_pad;
  View dst_plain;       // p == nullptr to skip (halo, if any, is pre-zeroed)
  int N, H, W;
};
void launch_img_prep(const ImgPrepParams& p, cudaStream_t s);

// L1 (cycle / identity) loss with gradient. pred: bf16 8-ch image, target: fp32 NHWC3.
//   loss_slot  += loss_scale * sum|t - p|         g = -grad_scale * sign(t - p)
struct L1LossParams {
  View pred;
  const float* target;
  float loss_scale, grad_scale;
  float* loss_slot;
  float* err_slot;      // optional second accumulator with err_scale (test-step MAE)
  float err_scale;
  View g;               // bf16 8-ch gradient (p == nullptr: loss only)
  int N, H, W;
};
void launch_l1_loss(const L1LossParams& p, cudaStream_t s);

// LSGAN loss on PatchGAN logits d (bf16, 1 valid channel at stride sw):
//   loss_slot += loss_scale * sum (target - d)^2     g = grad_scale * (d - target)
struct MseLossParams {
  View d;
  float target;
  float loss_scale, grad_scale;
  float* loss_slot;
  View g;               // bf16, 1 valid channel (p == nullptr: loss only)
  int N, H, W;
};
void launch_mse_loss(const MseLossParams& p, cudaStream_t s);

// Keras Adam on a flat fp32 buffer; also refreshes the bf16 compute copy.
struct AdamParams {
  float* w;
  float* m;
  float* v;
  const float* g;
  __nv_bfloat16* w_bf16;   // may be nullptr
  const int* step;         // device counter: number of updates already applied
  float lr, beta1, beta2, eps;
  float grad_scale;        // multiplies g (1.0 for plain SUM all-reduce)
  long long n;
};
void launch_adam(const AdamParams& p, cudaStream_t s);
void launch_step_increment(int* step, cudaStream_t s);

// out[i] = 0 where mask[i] == 0 (zero the gradients of structural padding weights)
void launch_mask_grad(float* g, const uint8_t* mask, long long n, cudaStream_t s);

// Horizontal tap gather of the row-GEMM form of the generator's 7x7 output conv:
//   out[n,h,w,co] = act(bias[co] + sum_s Q[flat(n,h,w) + s][s*4 + co]),  Q = [rows][32] bf16
struct TapGatherParams {
  const __nv_bfloat16* q;   // row-GEMM output, row = padded flat pixel, 32 columns (k*4 used)
  const float* bias;        // [3]
  View out;                 // bf16 8-channel image
  int act;
  int k;                    // taps per row (7)
  int Hp, Wp;               // padded dims of the flat domain
  int N, H, W;
};
void launch_tap_gather(const TapGatherParams& p, cudaStream_t s);

// FP8 weight quantisation + delayed-scaling bookkeeping (per FP8 layer l of a network):
//   w_amax:  amax_w[l] = max |w| over the layer's kernel segment
//   scales:  sa = pow2 scale from the previous step's activation amax (default when 0), sb from
//            amax_w; act_scale[l] = sa, w_scale[l] = sb, out_scale[l] = 1 / (sa * sb); amaxes reset
//   w_quant: w8 = sat_e4m3(w * w_scale[l])
constexpr int kMaxFp8Layers = 32;
struct Fp8Table {
  long long w_off[kMaxFp8Layers];     // element offset of the kernel segment in the flat buffers
  long long w_numel[kMaxFp8Layers];
  int num_layers;
  const __nv_bfloat16* w;             // bf16 compute copy of the weights (complete on every rank)
  uint8_t* w8;                        // E4M3 copy (same flat layout)
  float* act_amax;                    // [L]
  float* w_amax;                      // [L]
  float* act_scale;                   // [L]
  float* w_scale;                     // [L]
  float* out_scale;                   // [L]
  float default_act_scale;
};
void launch_fp8_w_amax(const Fp8Table& t, cudaStream_t s);
void launch_fp8_scales(const Fp8Table& t, cudaStream_t s);
void launch_fp8_w_quant(const Fp8Table& t, cudaStream_t s);

// PatchGAN output layer (512 -> 1) backward: SIMT kernels (GEMV-shaped).
struct D4BwdParams {
  View gd;              // bf16 gradient of the logits, 1 valid channel
  View a;               // bf16 input activation of the layer [N,H,W,C] (plain, OOB = zero pad)
  const __nv_bfloat16* w;  // bf16 weights [16 (cout padded)][16 taps][C]; row 0 used
  View dx;              // bf16 out [N,H,W,C] (p == nullptr to skip dgrad)
  float* dw;            // fp32 [16][16*C] grads (row 0), accumulated
  float* dbias;         // fp32 [1]
  int pad_before;       // TF SAME pad before (1 for k=4,s=1)
  int N, H, W, C;
};
void launch_d4_dgrad(const D4BwdParams& p, cudaStream_t s);
void launch_d4_wgrad(const D4BwdParams& p, cudaStream_t s);

}  // namespace b2
