// Implicit-GEMM convolution on tcgen05: shared parameter block.
#pragma once
#include "common.cuh"

namespace b2 {

constexpr int kMaxTaps = 52;   // 7x7 = 49 taps is the largest filter in the model
constexpr int kMaxAMaps = 4;   // stride-2 layers read 4 parity views of the same tensor

// One filter tap: which A tensor map to read, the coordinate offset of the
// 128-pixel tile for this tap, and the column (K offset) of its weights in B.
struct Tap {
  int16_t map;
  int16_t dn;
  int32_t dw;
  int32_t dh;
  int32_t bk;
};

constexpr int kMaxPhases = 4;   // sub-pixel phases of a transposed conv share one launch

struct ConvGemmParams {
  CUtensorMap tmA[kMaxAMaps];   // bf16 [.., C] views, box {64, bw, bh, 1}, SWIZZLE_128B
  CUtensorMap tmB;              // bf16 weights, box {64, BLOCK_N} (K-major) or {64, 64} (MN-major)
  Tap taps[kMaxTaps];
  int num_taps;
  int cin_blocks;               // Cin / 64 : k-blocks = num_taps * cin_blocks
  int tiles_w, tiles_h;         // tiles per image along w / h
  int bw, bh;                   // tile = bw x bh pixels, bw * bh == 128
  int w_limit;                  // rows with w coordinate >= w_limit are not stored (flat tails)
  __nv_bfloat16* out;           // NHWC-like: out[n*sn + h*sh + w*sw + c]
  long long out_sn, out_sh, out_sw;
  int cout;                     // valid output channels
  int num_phases;               // >= 1; tile t belongs to phase t / (m_tiles * n_tiles)
  int phase_tap_begin[kMaxPhases];
  int phase_tap_count[kMaxPhases];
  long long phase_out_off[kMaxPhases];   // element offset added to `out`
  const float* bias;            // [cout] or nullptr
  int act;                      // b2::Act, applied after bias
  float* stats;                 // [images][cout][2] (sum, sumsq) of the stored bf16 values, or nullptr
  const float* out_scale;       // device scalar multiplied into the accumulators (FP8 de-scaling), or nullptr
  int fp8;                      // operands are E4M3 bytes (128 channels per 128-byte K block)
};

void launch_conv_gemm(const ConvGemmParams& p, int block_n, bool b_mn_major, int num_images,
                      cudaStream_t stream);

}  // namespace b2
