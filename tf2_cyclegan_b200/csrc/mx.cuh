// MXFP8 (OCP microscaling) helpers: E4M3 elements with one E8M0 (power-of-two) scale per 32
// consecutive K elements, the operand format of tcgen05.mma.kind::mxf8f6f4.block_scale.
#pragma once
#include <cuda_fp8.h>

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

// One warp per pixel of a 256-channel tensor, 8 channels per lane: quantise this lane's values with the
// scale of their 32-channel block (4 adjacent lanes) and assemble the pixel's 8 scale bytes (every lane
// gets the record; lane 0 stores it). Shared by the IN-apply (forward operand) and IN-backward-apply
// (dgrad operand) twin writers.
__device__ __forceinline__ void mx_quant_pixel(const float (&v)[8], uint2& bytes, uint2& rec) {
  float bmax = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) bmax = fmaxf(bmax, fabsf(v[j]));
  bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, 1));
  bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, 2));
  const uint32_t sb = mx_scale_byte(bmax);
  const float inv = mx_inv_scale(sb);
  uint32_t o8[4];
  rec = make_uint2(0u, 0u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o8[j] = __nv_cvt_float2_to_fp8x2(make_float2(v[2 * j] * inv, v[2 * j + 1] * inv), __NV_SATFINITE, __NV_E4M3);
    rec.x |= __shfl_sync(0xffffffffu, sb, 4 * j) << (8 * j);
    rec.y |= __shfl_sync(0xffffffffu, sb, 16 + 4 * j) << (8 * j);
  }
  bytes = make_uint2(o8[0] | (o8[1] << 16), o8[2] | (o8[3] << 16));
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

// bf16 weights -> an E4M3 B matrix [rows][taps * kdim] (K-major) + its scales in the order above.
// Element (row n, tap t, k) is read from w[n * s_row + t * s_tap + k * s_k]: the forward operand of a conv
// stored as [cout][taps * cin] is (rows = cout, kdim = cin, s_row = taps * cin, s_tap = cin, s_k = 1), its
// dgrad operand W^T (rows = cin, kdim = cout, s_row = 1, s_tap = cin, s_k = taps * cin) — the K blocks of 32
// then run over OUTPUT channels, as the block scaling of the transposed product requires.
struct MxWeightQuantParams {
  const __nv_bfloat16* w;
  uint8_t* w8;
  uint8_t* sfb;
  int rows, taps, kdim;
  long long s_row, s_tap, s_k;
};
void launch_mx_weight_quant(const MxWeightQuantParams& p, cudaStream_t stream);

}  // namespace b2
