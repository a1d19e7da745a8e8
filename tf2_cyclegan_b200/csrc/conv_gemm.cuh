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

// Tap sharing: ONE halo tile of the activations (all pixels any tap of the tile touches) is
// loaded per 64-channel block and every tap reads it through a row-shifted UMMA descriptor
// (SWIZZLE_128B is a function of the absolute shared-memory address, so a K-major operand may
// start at any 128-byte row). Cuts the L2 -> SM activation traffic of a 3x3 conv from 9 tiles
// to 1.4 (8 x 16 pixel tiles) or 2.1 (flat 128-row tiles) per channel block.
struct TapShare {
  int enabled;
  int nbox;          // TMA boxes per halo tile (stacked along w; box_bytes % 1024 == 0 if > 1)
  int box_step;      // w-coordinate step between consecutive boxes
  int box_bytes;     // bytes per box = box rows x 128
  int pitch;         // halo rows per unit of dh: shift(tap) = (dh - origin_dh) * pitch + dw - origin_dw
  int origin_dw, origin_dh;   // coordinate offset of the halo origin relative to the tile origin
  int sbo;           // bytes between the 8-row groups of the M dimension inside the halo tile
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
  int h_limit;                  // rows with h coordinate >= h_limit are not stored (partial last tile row)
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
  TapShare ts;                  // optional (BLOCK_N == 256, cta_group::2, one phase, one A view)
  CUtensorMap tmTS;             // halo box view of the activations when ts.enabled
};

void launch_conv_gemm(const ConvGemmParams& p, int block_n, bool b_mn_major, int num_images,
                      cudaStream_t stream);

}  // namespace b2
