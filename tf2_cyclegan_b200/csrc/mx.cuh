// MXFP8 (OCP microscaling) helpers: E4M3 elements with one E8M0 (power-of-two) scale per 32
// consecutive K elements, the operand format of tcgen05.mma.kind::mxf8f6f4.block_scale.
#pragma once
#include "conv_gemm.cuh"

namespace b2 {

constexpr int kMxBlock = 32;       // elements per scale
constexpr int kMxSfRecord = 16;    // bytes per pixel in an activation scale buffer (one TMA-able record:
                                   // byte j = scale of channels [32 j, 32 j + 32), up to 512 channels)

// E8M0 byte (biased exponent, value 2^(byte - 127)) of the smallest power of two s with
// amax / s <= 448, the largest E4M3 magnitude. amax >= 0.
__device__ __forceinline__ uint32_t mx_scale_byte(float amax) {
  const uint32_t b = __float_as_uint(amax);
  // amax = m * 2^f, 1 <= m < 2: m * 2^8 <= 448 iff m <= 1.75 (mantissa field <= 0x600000)
  const int e = static_cast<int>(b >> 23) - 8 + ((b & 0x7FFFFFu) > 0x600000u ? 1 : 0);
  return static_cast<uint32_t>(min(max(e, 0), 254));
}
// 1 / scale as a float (exact: a power of two)
__device__ __forceinline__ float mx_inv_scale(uint32_t byte) {
  return __uint_as_float((254u - byte) << 23);
}

// Implicit-GEMM conv with MXFP8 operands (conv_gemm_mx.cu).
struct ConvGemmMxParams {
  ConvGemmParams g;     // tmA[0]: E4M3 activations, box {128 ch, bw, bh, 1}; tmB: E4M3 weights, box {128, 64};
                        // g.ts / g.tmTS: optional halo box of the activations (tap-shared CTA-pair kernel)
  CUtensorMap tmSFA;    // activation scale records [.., 16] uint8, box {16, bw, bh, 1}, no swizzle
  CUtensorMap tmSFTS;   // g.ts.enabled: halo box of the scale records (same pixels as g.tmTS, 16-byte inner box)
  const uint8_t* sfb;   // weight scales [n_tiles][cin_blocks][taps][32][4][4] (TMEM scale-factor order)
};
void launch_conv_gemm_mx(const ConvGemmMxParams& p, int num_images, cudaStream_t stream);
void set_mx_profile(unsigned long long* buf);   // stall attribution of the tap-shared kernel (tools/kernel_bench.py)

// bf16 [cout][taps * cin] weights -> E4M3 bytes (same indexing) + scales in the order above
struct MxWeightQuantParams {
  const __nv_bfloat16* w;
  uint8_t* w8;
  uint8_t* sfb;
  int cout, taps, cin;
};
void launch_mx_weight_quant(const MxWeightQuantParams& p, cudaStream_t stream);

}  // namespace b2
