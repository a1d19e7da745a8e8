// Weight-gradient GEMM (split-K, MN-major operands): shared parameter block.
#pragma once
#include "conv_gemm.cuh"

namespace b2 {

struct WgradParams {
  CUtensorMap tmDY[kMaxAMaps];  // bf16 gradient views [.., Cout], box {64, kbw, kbh, 1}; tap.dn selects one
  CUtensorMap tmA[kMaxAMaps];   // bf16 activation views [.., Cin], box {64, kbw, kbh, 1}
  Tap taps[kMaxTaps];           // same tap table as the forward conv (bk = dW column)
  int num_taps;
  int ktiles_w, ktiles_h;       // K tiles (64 pixels each) per image
  int kbw, kbh;                 // kbw * kbh == 64
  int num_images;
  int cin, cout;
  float* dw;                    // fp32 [cout][ldw], column = tap.bk + ci
  long long ldw;
};

void launch_wgrad_gemm(const WgradParams& p, int block_n, int ksplit, cudaStream_t stream);

}  // namespace b2

namespace b2 {

// Tap-shared wgrad of a 3x3 stride-1 reflect-padded conv (the residual-block layers): a CTA pair
// owns 256 output channels x 128 input channels and ALL THREE horizontal taps of one kernel row.
// The activation row (64 + 2 padded pixels) is loaded once and read through row-shifted UMMA
// descriptors (+128 B per tap), dY once; three accumulators live in TMEM (3 x 128 columns).
struct WgradRowParams {
  CUtensorMap tmDY;    // dY interior view [.., Cout], box {64, 64, 1, 1}
  CUtensorMap tmA;     // padded activation view [.., Cin], box {64, 66, 1, 1}
  int tiles_w;         // 64-pixel K tiles per image row (W / 64)
  int H;               // image rows
  int num_images;
  int cin, cout;       // cin % 128 == 0, cout % 256 == 0
  float* dw;           // fp32 [cout][9 * cin]
  long long ldw;
};
void launch_wgrad_row(const WgradRowParams& p, int ksplit, cudaStream_t stream);

}  // namespace b2
