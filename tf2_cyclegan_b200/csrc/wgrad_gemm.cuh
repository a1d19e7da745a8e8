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

}  // namespace b2
