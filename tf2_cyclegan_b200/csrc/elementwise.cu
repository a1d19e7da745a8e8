// Bandwidth-bound kernels around the GEMMs (see elementwise.cuh).
#include <stdlib.h>

#include <cuda_fp8.h>

#include "elementwise.cuh"
#include "mx.cuh"

namespace b2 {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ const __nv_bfloat16* vptr(const View& v, int n, int h, int w) {
  return reinterpret_cast<const __nv_bfloat16*>(v.p) + n * v.sn + h * v.sh + w * v.sw;
}
__device__ __forceinline__ __nv_bfloat16* vptr_mut(const View& v, int n, int h, int w) {
  return reinterpret_cast<__nv_bfloat16*>(v.p) + n * v.sn + h * v.sh + w * v.sw;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = unpack_bf16x2(w[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) = pack8(f);
}

// reflected halo coordinates of interior coordinate i (size n, pad p): up to 2 values incl. i
__device__ __forceinline__ int reflect_set(int i, int n, int p, int (&out)[2]) {
  out[0] = i;
  int cnt = 1;
  if (p > 0) {
    if (i >= 1 && i <= p) out[cnt++] = -i;
    else if (i >= n - 1 - p && i <= n - 2) out[cnt++] = 2 * (n - 1) - i;
  }
  return cnt;
}

// linear pixel index -> (n, h, w) with 32-bit unsigned arithmetic (every tensor here has
// far fewer than 2^31 pixels; 64-bit div / mod costs ~5x more instructions)
__device__ __forceinline__ void unravel(uint32_t i, int W, int H, int& n, int& h, int& w) {
  const uint32_t r = i / static_cast<uint32_t>(W);
  w = static_cast<int>(i - r * static_cast<uint32_t>(W));
  const uint32_t q = r / static_cast<uint32_t>(H);
  h = static_cast<int>(r - q * static_cast<uint32_t>(H));
  n = static_cast<int>(q);
}

__device__ __forceinline__ float act_grad_from_pre(float y, float slope) {
  return y > 0.f ? 1.f : slope;
}

// gradient gather with optional reflect-fold (sum of the halo positions that mirror (h, w))
__device__ __forceinline__ void gather_g(const View& g1, int fold, const View& g2, int n, int h,
                                         int w, int H, int W, int c, float (&g)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = 0.f;
  int hs[2], ws[2];
  const int nh = reflect_set(h, H, fold, hs);
  const int nw = reflect_set(w, W, fold, ws);
  for (int a = 0; a < nh; ++a)
    for (int b = 0; b < nw; ++b) {
      float t[8];
      load8(vptr(g1, n, hs[a], ws[b]) + c, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += t[j];
    }
  if (g2.p != nullptr) {
    float t[8];
    load8(vptr(g2, n, h, w) + c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] += t[j];
  }
}

// ---------------------------------------------------------- InstanceNorm kernels
// Shared structure of in_apply / in_bwd_reduce / in_bwd_apply. grid: (pixel chunks, N). A thread
// owns ONE 8-channel group (16-byte vectors; its per-channel coefficients live in registers) and
// walks the pixels ln, ln + lanes, ... of the block's chunk, U pixels (= U independent 16-byte
// loads per tensor) in flight. Pixel -> (h, w) is one multiply-high with a host-computed
// reciprocal (no division, no per-slot state); every in-image offset is 32-bit arithmetic; the
// 64-bit image base is added once per block. Feature switches (activation kind, residual input,
// reflect halo, second gradient, FP8 twin) are template parameters so that the unused paths cost
// no instructions or registers: these kernels are issue-bound long before they are memory-bound
// when written generically (round 1: ~500 SASS instructions per 16-byte vector, 34-50 % of HBM).
constexpr int ACTK_NONE = 0, ACTK_RELU = 1, ACTK_SLOPE = 2;

struct V32 {   // View with 32-bit in-image strides, image base resolved
  const __nv_bfloat16* p;
  int sh, sw;
};
__device__ __forceinline__ V32 v32(const View& v, int n, int c) {
  V32 r;
  r.p = reinterpret_cast<const __nv_bfloat16*>(v.p) + static_cast<long long>(n) * v.sn + c;
  r.sh = static_cast<int>(v.sh);
  r.sw = static_cast<int>(v.sw);
  return r;
}
__device__ __forceinline__ uint4 ldv(const V32& v, int h, int w) {
  return __ldg(reinterpret_cast<const uint4*>(v.p + (h * v.sh + w * v.sw)));
}
__device__ __forceinline__ void stv(const V32& v, int h, int w, const uint4& u) {
  *reinterpret_cast<uint4*>(const_cast<__nv_bfloat16*>(v.p) + (h * v.sh + w * v.sw)) = u;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  // bf16 -> fp32 is a 16-bit shift: low half << 16, high half masked
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ void load_f8(const float* p, float (&f)[8]) {   // 32-byte aligned
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
// (mean, rstd) of the thread's 8 channels from the forward statistics [.. 2 * c ..] = (sum, sumsq)
__device__ __forceinline__ void mean_rstd8(const float* st, float inv_count, float eps, float (&mean)[8],
                                           float (&rstd)[8]) {
  float a[8], b[8];
  load_f8(st, a);
  load_f8(st + 8, b);
  const float s[16] = {a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7]};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mean[j] = s[2 * j] * inv_count;
    rstd[j] = rsqrtf(fmaxf(s[2 * j + 1] * inv_count - mean[j] * mean[j], 0.f) + eps);
  }
}
template <int ACT>
__device__ __forceinline__ float actk(float v, float slope) {
  if constexpr (ACT == ACTK_RELU) return fmaxf(v, 0.f);
  if constexpr (ACT == ACTK_SLOPE) return v > 0.f ? v : v * slope;
  return v;
}

// ------------------------------------------------------------------ in_apply
// dst = act(a * x + b) (+ skip), written into the (reflect-haloed) input buffer of the next conv
// F8: 0 = no 8-bit twin, 1 = E4M3 with one delayed scale per tensor, 2 = MXFP8 (E4M3 + one E8M0 scale per
// 32 channels, C == 256: one warp per pixel)
template <int ACT, bool SKIP, bool HALO, int F8>
__global__ void __launch_bounds__(kThreads, (SKIP || F8 != 0) ? 3 : 4) in_apply_kernel(const InApplyParams p, int pix_per_block,
                                                              uint32_t rcp_w) {
  constexpr int U = 4;
  const int c8 = p.C >> 3;
  const int lanes = kThreads / c8;
  const int c = (threadIdx.x % c8) * 8;
  const int ln = threadIdx.x / c8;
  const int n = blockIdx.y;
  const int HW = p.H * p.W;
  const int pix0 = blockIdx.x * pix_per_block;
  const int pix1 = min(HW, pix0 + pix_per_block);
  float ka[8], kb[8];
  if (p.stats != nullptr) {
    float mean[8], rstd[8], ga[8], be[8];
    mean_rstd8(p.stats + (static_cast<long long>(n) * p.C + c) * 2, p.inv_count, p.eps, mean, rstd);
    load_f8(p.gamma + c, ga);
    load_f8(p.beta + c, be);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ka[j] = rstd[j] * ga[j];
      kb[j] = be[j] - mean[j] * ka[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) { ka[j] = 1.f; kb[j] = 0.f; }
  }
  const V32 X = v32(p.x, n, c), S = v32(p.skip, n, c), D = v32(p.dst, n, c);
  constexpr bool HAS8 = F8 != 0;
  uint8_t* d8 = nullptr;
  uint8_t* sf = nullptr;
  int d8sh = 0, d8sw = 0, sfsh = 0, sfsw = 0;
  float s8 = 1.f, amax = 0.f;
  if constexpr (HAS8) {
    d8 = reinterpret_cast<uint8_t*>(p.dst8.p) + static_cast<long long>(n) * p.dst8.sn + c;
    d8sh = static_cast<int>(p.dst8.sh);
    d8sw = static_cast<int>(p.dst8.sw);
    if constexpr (F8 == 1) s8 = __ldg(p.scale8);
    if constexpr (F8 == 2) {
      sf = reinterpret_cast<uint8_t*>(p.sf8.p) + static_cast<long long>(n) * p.sf8.sn;
      sfsh = static_cast<int>(p.sf8.sh);
      sfsw = static_cast<int>(p.sf8.sw);
    }
  }
  const float slope = act_slope(p.act);
  const int H = p.H, W = p.W, halo = p.halo;
  for (int base = pix0 + ln; base < pix1; base += lanes * U) {
    uint4 xv[U], sv[U];
    int hh[U], ww[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = base + u * lanes;
      hh[u] = static_cast<int>(__umulhi(static_cast<uint32_t>(pix), rcp_w));
      ww[u] = pix - hh[u] * W;
      if (pix < pix1) {
        xv[u] = ldv(X, hh[u], ww[u]);
        if constexpr (SKIP) sv[u] = ldv(S, hh[u], ww[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * lanes >= pix1) continue;
      float x[8], v[8];
      unpack8(xv[u], x);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = actk<ACT>(fmaf(x[j], ka[j], kb[j]), slope);
      if constexpr (SKIP) {
        float s[8];
        unpack8(sv[u], s);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += s[j];
      }
      const uint4 pk = pack8(v);
      const int h = hh[u], w = ww[u];
      stv(D, h, w, pk);
      uint2 pk8 = make_uint2(0u, 0u);
      uint2 sfrec = make_uint2(0u, 0u);   // MX: the pixel's 8 scale bytes (lane 0)
      if constexpr (HAS8) {
        uint16_t o8[4];
        if constexpr (F8 == 2) {
          // the 32-channel block = 4 adjacent lanes; the whole warp works on this pixel
          float bmax = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) bmax = fmaxf(bmax, fabsf(v[j]));
          bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, 1));
          bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, 2));
          const uint32_t sb = mx_scale_byte(bmax);
          s8 = mx_inv_scale(sb);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            sfrec.x |= __shfl_sync(0xffffffffu, sb, 4 * j) << (8 * j);
            sfrec.y |= __shfl_sync(0xffffffffu, sb, 16 + 4 * j) << (8 * j);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (F8 == 1) amax = fmaxf(amax, fmaxf(fabsf(v[2 * j]), fabsf(v[2 * j + 1])));
          o8[j] = __nv_cvt_float2_to_fp8x2(make_float2(v[2 * j] * s8, v[2 * j + 1] * s8), __NV_SATFINITE, __NV_E4M3);
        }
        pk8 = make_uint2(static_cast<uint32_t>(o8[0]) | (static_cast<uint32_t>(o8[1]) << 16),
                         static_cast<uint32_t>(o8[2]) | (static_cast<uint32_t>(o8[3]) << 16));
        *reinterpret_cast<uint2*>(d8 + (h * d8sh + w * d8sw)) = pk8;
        if constexpr (F8 == 2) {
          if ((threadIdx.x & 31) == 0) *reinterpret_cast<uint2*>(sf + (h * sfsh + w * sfsw)) = sfrec;
        }
      }
      if constexpr (HALO) {
        // mirror images of a border pixel: row hm and / or column wm (reflect: -i for 1 <= i <= halo,
        // 2 (n - 1) - i for n - 1 - halo <= i <= n - 2)
        const int hm = (h >= 1 && h <= halo) ? -h : ((h >= H - 1 - halo && h <= H - 2) ? 2 * (H - 1) - h : h);
        const int wm = (w >= 1 && w <= halo) ? -w : ((w >= W - 1 - halo && w <= W - 2) ? 2 * (W - 1) - w : w);
        if (hm != h) {
          stv(D, hm, w, pk);
          if constexpr (HAS8) *reinterpret_cast<uint2*>(d8 + (hm * d8sh + w * d8sw)) = pk8;
          if constexpr (F8 == 2) { if ((threadIdx.x & 31) == 0) *reinterpret_cast<uint2*>(sf + (hm * sfsh + w * sfsw)) = sfrec; }
        }
        if (wm != w) {
          stv(D, h, wm, pk);
          if constexpr (HAS8) *reinterpret_cast<uint2*>(d8 + (h * d8sh + wm * d8sw)) = pk8;
          if constexpr (F8 == 2) { if ((threadIdx.x & 31) == 0) *reinterpret_cast<uint2*>(sf + (h * sfsh + wm * sfsw)) = sfrec; }
          if (hm != h) {
            stv(D, hm, wm, pk);
            if constexpr (HAS8) *reinterpret_cast<uint2*>(d8 + (hm * d8sh + wm * d8sw)) = pk8;
            if constexpr (F8 == 2) { if ((threadIdx.x & 31) == 0) *reinterpret_cast<uint2*>(sf + (hm * sfsh + wm * sfsw)) = sfrec; }
          }
        }
      }
    }
  }
  if constexpr (F8 == 1) {   // delayed-scaling statistics: one atomic per warp (values are >= 0, so int order == float order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(p.amax8), __float_as_int(amax));
  }
}

// ------------------------------------------------------------ in_bwd_reduce
// dy = (fold(g1) + g2) * act'(y); bstats[n][c] += (sum dy, sum dy * xhat). The sums run over the raw
// x (sum dy * x); xhat = (x - mean) * rstd is applied once per block at the end.
template <int ACT, bool FOLD, bool G2, bool HAS_IN>
__global__ void __launch_bounds__(kThreads, 3) in_bwd_reduce_kernel(const InBwdReduceParams p, int pix_per_block,
                                                                   uint32_t rcp_w) {
  constexpr int U = 2;
  extern __shared__ float red[];   // [lanes][C][2]
  const int c8 = p.C >> 3;
  const int lanes = kThreads / c8;
  const int c = (threadIdx.x % c8) * 8;
  const int ln = threadIdx.x / c8;
  const int n = blockIdx.y;
  const int HW = p.H * p.W;
  const int pix0 = blockIdx.x * pix_per_block;
  const int pix1 = min(HW, pix0 + pix_per_block);
  float ka[8], kb[8];   // y = ka * x + kb: only its sign is needed (act')
  if constexpr (HAS_IN && ACT != ACTK_NONE) {
    float mean[8], rstd[8], ga[8], be[8];
    mean_rstd8(p.stats + (static_cast<long long>(n) * p.C + c) * 2, p.inv_count, p.eps, mean, rstd);
    load_f8(p.gamma + c, ga);
    load_f8(p.beta + c, be);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ka[j] = rstd[j] * ga[j];
      kb[j] = be[j] - mean[j] * ka[j];
    }
  }
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  const V32 G1 = v32(p.g1, n, c), G2v = v32(p.g2, n, c), X = v32(p.x, n, c), DY = v32(p.dy, n, c);
  const float slope = act_slope(p.act);
  const int H = p.H, W = p.W, fold = p.fold1;
  for (int base = pix0 + ln; base < pix1; base += lanes * U) {
    uint4 gv[U], g2v[U], xv[U];
    int hh[U], ww[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = base + u * lanes;
      hh[u] = static_cast<int>(__umulhi(static_cast<uint32_t>(pix), rcp_w));
      ww[u] = pix - hh[u] * W;
      if (pix < pix1) {
        gv[u] = ldv(G1, hh[u], ww[u]);
        if constexpr (G2) g2v[u] = ldv(G2v, hh[u], ww[u]);
        if constexpr (HAS_IN) xv[u] = ldv(X, hh[u], ww[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * lanes >= pix1) continue;
      const int h = hh[u], w = ww[u];
      float g[8];
      unpack8(gv[u], g);
      if constexpr (G2) {
        float t[8];
        unpack8(g2v[u], t);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] += t[j];
      }
      if constexpr (FOLD) {   // border pixels also collect the gradient of their mirror images in g1's halo
        const int hm = (h >= 1 && h <= fold) ? -h : ((h >= H - 1 - fold && h <= H - 2) ? 2 * (H - 1) - h : h);
        const int wm = (w >= 1 && w <= fold) ? -w : ((w >= W - 1 - fold && w <= W - 2) ? 2 * (W - 1) - w : w);
        if (hm != h || wm != w) {
          float t[8];
          if (hm != h) {
            unpack8(ldv(G1, hm, w), t);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] += t[j];
          }
          if (wm != w) {
            unpack8(ldv(G1, h, wm), t);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] += t[j];
            if (hm != h) {
              unpack8(ldv(G1, hm, wm), t);
#pragma unroll
              for (int j = 0; j < 8; ++j) g[j] += t[j];
            }
          }
        }
      }
      if constexpr (HAS_IN) {
        float x[8];
        unpack8(xv[u], x);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if constexpr (ACT != ACTK_NONE) g[j] = fmaf(x[j], ka[j], kb[j]) > 0.f ? g[j] : g[j] * slope;
        }
        const uint4 pk = pack8(g);
        float gq[8];
        unpack8(pk, gq);   // statistics of the STORED (bf16) value: in_bwd_apply reads that
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += gq[j];
          s2[j] = fmaf(gq[j], x[j], s2[j]);
        }
        stv(DY, h, w, pk);
      } else {
        stv(DY, h, w, pack8(g));
      }
    }
  }
  if (!HAS_IN || p.bstats == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[(ln * p.C + c + j) * 2 + 0] = s1[j];
    red[(ln * p.C + c + j) * 2 + 1] = s2[j];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < p.C; ch += kThreads) {
    float a1 = 0.f, a2 = 0.f;
    for (int l = 0; l < lanes; ++l) {
      const float2 t = *reinterpret_cast<const float2*>(&red[(l * p.C + ch) * 2]);
      a1 += t.x;
      a2 += t.y;
    }
    const float* st = p.stats + (static_cast<long long>(n) * p.C + ch) * 2;
    const float mean = st[0] * p.inv_count;
    const float rstd = rsqrtf(fmaxf(st[1] * p.inv_count - mean * mean, 0.f) + p.eps);
    float* bs = p.bstats + (static_cast<long long>(n) * p.C + ch) * 2;
    atomicAdd(bs, a1);
    atomicAdd(bs + 1, rstd * (a2 - mean * a1));
  }
}

// ------------------------------------------------------------- in_bwd_apply
// dx = A*dy + B*x + C per channel (A = gamma*rstd, B = -gamma*rstd^2*m2, C = gamma*rstd*(mean*rstd*m2 - m1))
__global__ void __launch_bounds__(kThreads, 3) in_bwd_apply_kernel(const InBwdApplyParams p, int pix_per_block,
                                                                  uint32_t rcp_w) {
  constexpr int U = 4;
  const int c8 = p.C >> 3;
  const int lanes = kThreads / c8;
  const int c = (threadIdx.x % c8) * 8;
  const int ln = threadIdx.x / c8;
  const int n = blockIdx.y;
  const int HW = p.H * p.W;
  const int pix0 = blockIdx.x * pix_per_block;
  const int pix1 = min(HW, pix0 + pix_per_block);
  if (blockIdx.x == 0 && blockIdx.y == 0 && p.dgamma != nullptr) {   // parameter gradients: sum over images
    for (int cc = threadIdx.x; cc < p.C; cc += blockDim.x) {
      float dg = 0.f, db = 0.f;
      for (int nn = 0; nn < p.N; ++nn) {
        const float* bs = p.bstats + (static_cast<long long>(nn) * p.C + cc) * 2;
        db += bs[0];
        dg += bs[1];
      }
      p.dgamma[cc] += dg;
      p.dbeta[cc] += db;
    }
  }
  float kA[8], kB[8], kC[8];
  {
    float mean[8], rstd[8], ga[8], b0[8], b1[8];
    mean_rstd8(p.stats + (static_cast<long long>(n) * p.C + c) * 2, p.inv_count, p.eps, mean, rstd);
    load_f8(p.gamma + c, ga);
    load_f8(p.bstats + (static_cast<long long>(n) * p.C + c) * 2, b0);
    load_f8(p.bstats + (static_cast<long long>(n) * p.C + c) * 2 + 8, b1);
    const float bs[16] = {b0[0], b0[1], b0[2], b0[3], b0[4], b0[5], b0[6], b0[7],
                          b1[0], b1[1], b1[2], b1[3], b1[4], b1[5], b1[6], b1[7]};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float m1 = bs[2 * j] * p.inv_count, m2 = bs[2 * j + 1] * p.inv_count;
      const float gr = ga[j] * rstd[j];
      kA[j] = gr;
      kB[j] = -gr * rstd[j] * m2;
      kC[j] = gr * (mean[j] * rstd[j] * m2 - m1);
    }
  }
  const V32 DYv = v32(p.dy, n, c), X = v32(p.x, n, c), DX = v32(p.dx, n, c);
  const int W = p.W;
  for (int base = pix0 + ln; base < pix1; base += lanes * U) {
    uint4 dv[U], xv[U];
    int hh[U], ww[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = base + u * lanes;
      hh[u] = static_cast<int>(__umulhi(static_cast<uint32_t>(pix), rcp_w));
      ww[u] = pix - hh[u] * W;
      if (pix < pix1) {
        dv[u] = ldv(DYv, hh[u], ww[u]);
        xv[u] = ldv(X, hh[u], ww[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * lanes >= pix1) continue;
      float d[8], x[8], o[8];
      unpack8(dv[u], d);
      unpack8(xv[u], x);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(kA[j], d[j], fmaf(kB[j], x[j], kC[j]));
      stv(DX, hh[u], ww[u], pack8(o));
    }
  }
}

// ------------------------------------------------ streamed InstanceNorm kernels
// The same three kernels with their INPUT tensors streamed through a shared-memory ring by the
// bulk-copy engine: one producer warp issues 1-D cp.async.bulk copies (pixel rows are contiguous
// in every NHWC view, so a stage of kSpx pixels is one or two copies per tensor) into a ring of
// kStages stages guarded by full / empty mbarriers; eight consumer warps read their 16-byte
// vectors from the ring (a warp reads 512 consecutive bytes: conflict-free), do the arithmetic and
// store straight to global memory. Loads in flight per SM = ring size (up to ~190 KB with two
// resident blocks) instead of what fits in registers, and they stay in flight while the consumers
// compute: the direct kernels above spend most of their ~15 us in three or four serialised
// load -> compute -> store rounds (latency bound, 12 us at zero bytes), these are bandwidth bound.
constexpr int kConsumers = 256;                // 8 consumer warps
constexpr int kStreamThreads = kConsumers + 32;   // + 1 producer warp
constexpr int kPxT = 2;                        // pixels per consumer thread per stage
constexpr int kStageBytes = kConsumers * 16 * kPxT;   // per tensor and stage: 8 KB

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct StreamGeom {
  int W, C, pix0, pix1, spx;   // chunk [pix0, pix1) of one image, spx pixels per stage
  uint32_t rcp_w;
};

// producer: copy pixels [pix, pix + npix) of one image view (pixel stride == C) to `dst`, row by row
__device__ __forceinline__ void stream_issue(uint8_t* dst, const V32& v, int pix, int npix, const StreamGeom& g,
                                             uint64_t* bar) {
  while (npix > 0) {
    const int h = static_cast<int>(__umulhi(static_cast<uint32_t>(pix), g.rcp_w));
    const int w = pix - h * g.W;
    const int run = min(npix, g.W - w);
    const uint32_t bytes = static_cast<uint32_t>(run) * g.C * 2;
    bulk_g2s(dst, v.p + (h * v.sh + w * g.C), bytes, bar);
    dst += bytes;
    pix += run;
    npix -= run;
  }
}

// Runs the ring. `views`: the NT streamed tensors (channel offset 0). Consumers call
// f(pix, h, w, vec[NT]) for each of their pixels; everything else (stores, statistics) is in f.
// smem: [kStages][NT][kStageBytes] data, then 2 * kStages mbarriers.
template <int NT, int STAGES, class F>
__device__ __forceinline__ void stream_run(uint8_t* smem, const V32 (&views)[NT], const StreamGeom& g, int c, int ln,
                                           int lanes, F&& f) {
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * NT * kStageBytes);
  uint64_t* empty = full + STAGES;
  const int total = (g.pix1 - g.pix0 + g.spx - 1) / g.spx;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumers / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x >= kConsumers) {
    // ===== producer warp (one elected lane issues) =====
    if (threadIdx.x == kConsumers) {
      for (int it = 0; it < total; ++it) {
        const int s = it % STAGES, fill = it / STAGES;
        if (fill > 0) mbar_wait(&empty[s], (fill - 1) & 1);
        const int pix = g.pix0 + it * g.spx;
        const int npix = min(g.spx, g.pix1 - pix);
        mbar_arrive_expect_tx(&full[s], static_cast<uint32_t>(npix) * g.C * 2 * NT);
#pragma unroll
        for (int t = 0; t < NT; ++t)
          stream_issue(smem + (s * NT + t) * kStageBytes, views[t], pix, npix, g, &full[s]);
      }
    }
    return;
  }
  // ===== consumers =====
  for (int it = 0; it < total; ++it) {
    const int s = it % STAGES;
    mbar_wait(&full[s], (it / STAGES) & 1);
    const int pstage = g.pix0 + it * g.spx;
    const int npix = min(g.spx, g.pix1 - pstage);
    uint4 v[kPxT][NT];
#pragma unroll
    for (int k = 0; k < kPxT; ++k) {
      const int lp = ln + k * lanes;
      if (lp < npix) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
          v[k][t] = *reinterpret_cast<const uint4*>(smem + (s * NT + t) * kStageBytes + (lp * g.C + c) * 2);
      }
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);   // the stage is in registers: hand it back
#pragma unroll
    for (int k = 0; k < kPxT; ++k) {
      const int lp = ln + k * lanes;
      if (lp < npix) {
        const int pix = pstage + lp;
        const int h = static_cast<int>(__umulhi(static_cast<uint32_t>(pix), g.rcp_w));
        f(pix, h, pix - h * g.W, v[k]);
      }
    }
  }
}

// MX: also write the MXFP8 twin (dst8 bytes + sf8 scale records; C == 256: one consumer warp per pixel)
template <int ACT, bool SKIP, bool HALO, bool MX>
__global__ void __launch_bounds__(kStreamThreads, 2) in_apply_stream_kernel(const InApplyParams p, int pix_per_block,
                                                                           uint32_t rcp_w) {
  extern __shared__ __align__(128) uint8_t ring[];
  constexpr int NT = SKIP ? 2 : 1;
  constexpr int STAGES = SKIP ? 6 : 8;
  const int c8 = p.C >> 3;
  const int lanes = kConsumers / c8;
  const int tid = threadIdx.x < kConsumers ? threadIdx.x : 0;
  const int c = (tid % c8) * 8;
  const int ln = tid / c8;
  const int n = blockIdx.y;
  StreamGeom g;
  g.W = p.W; g.C = p.C; g.rcp_w = rcp_w; g.spx = lanes * kPxT;
  g.pix0 = blockIdx.x * pix_per_block;
  g.pix1 = min(p.H * p.W, g.pix0 + pix_per_block);
  V32 views[NT];
  views[0] = v32(p.x, n, 0);
  if constexpr (SKIP) views[1] = v32(p.skip, n, 0);
  float ka[8], kb[8];
  if (threadIdx.x < kConsumers) {   // overlaps with the producer's first copies
    if (p.stats != nullptr) {
      float mean[8], rstd[8], ga[8], be[8];
      mean_rstd8(p.stats + (static_cast<long long>(n) * p.C + c) * 2, p.inv_count, p.eps, mean, rstd);
      load_f8(p.gamma + c, ga);
      load_f8(p.beta + c, be);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ka[j] = rstd[j] * ga[j];
        kb[j] = be[j] - mean[j] * ka[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { ka[j] = 1.f; kb[j] = 0.f; }
    }
  }
  const V32 D = v32(p.dst, n, c);
  const float slope = act_slope(p.act);
  const int H = p.H, W = p.W, halo = p.halo;
  uint8_t* d8 = nullptr;
  uint8_t* sf = nullptr;
  int d8sh = 0, d8sw = 0, sfsh = 0, sfsw = 0;
  if constexpr (MX) {
    d8 = reinterpret_cast<uint8_t*>(p.dst8.p) + static_cast<long long>(n) * p.dst8.sn + c;
    d8sh = static_cast<int>(p.dst8.sh);
    d8sw = static_cast<int>(p.dst8.sw);
    sf = reinterpret_cast<uint8_t*>(p.sf8.p) + static_cast<long long>(n) * p.sf8.sn;
    sfsh = static_cast<int>(p.sf8.sh);
    sfsw = static_cast<int>(p.sf8.sw);
  }
  const bool lane0 = (threadIdx.x & 31) == 0;
  stream_run<NT, STAGES>(ring, views, g, c, ln, lanes, [&](int, int h, int w, const uint4 (&v)[NT]) {
    float x[8], o[8];
    unpack8(v[0], x);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = actk<ACT>(fmaf(x[j], ka[j], kb[j]), slope);
    if constexpr (SKIP) {
      float s[8];
      unpack8(v[1], s);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += s[j];
    }
    const uint4 pk = pack8(o);
    stv(D, h, w, pk);
    uint2 pk8 = make_uint2(0u, 0u), sfrec = make_uint2(0u, 0u);
    if constexpr (MX) {
      mx_quant_pixel(o, pk8, sfrec);   // this warp == this pixel (C == 256); quantises the fp32 values
      *reinterpret_cast<uint2*>(d8 + (h * d8sh + w * d8sw)) = pk8;
      if (lane0) *reinterpret_cast<uint2*>(sf + (h * sfsh + w * sfsw)) = sfrec;
    }
    if constexpr (HALO) {
      const int hm = (h >= 1 && h <= halo) ? -h : ((h >= H - 1 - halo && h <= H - 2) ? 2 * (H - 1) - h : h);
      const int wm = (w >= 1 && w <= halo) ? -w : ((w >= W - 1 - halo && w <= W - 2) ? 2 * (W - 1) - w : w);
      auto mirror = [&](int hh, int ww) {
        stv(D, hh, ww, pk);
        if constexpr (MX) {
          *reinterpret_cast<uint2*>(d8 + (hh * d8sh + ww * d8sw)) = pk8;
          if (lane0) *reinterpret_cast<uint2*>(sf + (hh * sfsh + ww * sfsw)) = sfrec;
        }
      };
      if (hm != h) mirror(hm, w);
      if (wm != w) {
        mirror(h, wm);
        if (hm != h) mirror(hm, wm);
      }
    }
  });
}

template <int ACT, bool FOLD, bool G2>
__global__ void __launch_bounds__(kStreamThreads, 2) in_bwd_reduce_stream_kernel(const InBwdReduceParams p,
                                                                                int pix_per_block, uint32_t rcp_w) {
  extern __shared__ __align__(128) uint8_t ring[];
  constexpr int NT = G2 ? 3 : 2;       // g1, x, (g2)
  constexpr int STAGES = G2 ? 4 : 6;
  const int c8 = p.C >> 3;
  const int lanes = kConsumers / c8;
  const int tid = threadIdx.x < kConsumers ? threadIdx.x : 0;
  const int c = (tid % c8) * 8;
  const int ln = tid / c8;
  const int n = blockIdx.y;
  StreamGeom g;
  g.W = p.W; g.C = p.C; g.rcp_w = rcp_w; g.spx = lanes * kPxT;
  g.pix0 = blockIdx.x * pix_per_block;
  g.pix1 = min(p.H * p.W, g.pix0 + pix_per_block);
  V32 views[NT];
  views[0] = v32(p.g1, n, 0);
  views[1] = v32(p.x, n, 0);
  if constexpr (G2) views[2] = v32(p.g2, n, 0);
  float ka[8], kb[8];
  if constexpr (ACT != ACTK_NONE) {
    if (threadIdx.x < kConsumers) {
      float mean[8], rstd[8], ga[8], be[8];
      mean_rstd8(p.stats + (static_cast<long long>(n) * p.C + c) * 2, p.inv_count, p.eps, mean, rstd);
      load_f8(p.gamma + c, ga);
      load_f8(p.beta + c, be);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ka[j] = rstd[j] * ga[j];
        kb[j] = be[j] - mean[j] * ka[j];
      }
    }
  }
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  const V32 G1 = v32(p.g1, n, c), DY = v32(p.dy, n, c);
  const float slope = act_slope(p.act);
  const int H = p.H, W = p.W, fold = p.fold1;
  stream_run<NT, STAGES>(ring, views, g, c, ln, lanes, [&](int, int h, int w, const uint4 (&v)[NT]) {
    float gr[8], x[8];
    unpack8(v[0], gr);
    unpack8(v[1], x);
    if constexpr (G2) {
      float t[8];
      unpack8(v[2], t);
#pragma unroll
      for (int j = 0; j < 8; ++j) gr[j] += t[j];
    }
    if constexpr (FOLD) {   // border pixels also collect the gradient of their mirror images in g1's halo
      const int hm = (h >= 1 && h <= fold) ? -h : ((h >= H - 1 - fold && h <= H - 2) ? 2 * (H - 1) - h : h);
      const int wm = (w >= 1 && w <= fold) ? -w : ((w >= W - 1 - fold && w <= W - 2) ? 2 * (W - 1) - w : w);
      if (hm != h || wm != w) {
        float t[8];
        if (hm != h) {
          unpack8(ldv(G1, hm, w), t);
#pragma unroll
          for (int j = 0; j < 8; ++j) gr[j] += t[j];
        }
        if (wm != w) {
          unpack8(ldv(G1, h, wm), t);
#pragma unroll
          for (int j = 0; j < 8; ++j) gr[j] += t[j];
          if (hm != h) {
            unpack8(ldv(G1, hm, wm), t);
#pragma unroll
            for (int j = 0; j < 8; ++j) gr[j] += t[j];
          }
        }
      }
    }
    if constexpr (ACT != ACTK_NONE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) gr[j] = fmaf(x[j], ka[j], kb[j]) > 0.f ? gr[j] : gr[j] * slope;
    }
    const uint4 pk = pack8(gr);
    float gq[8];
    unpack8(pk, gq);   // statistics of the STORED (bf16) value: in_bwd_apply reads that
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1[j] += gq[j];
      s2[j] = fmaf(gq[j], x[j], s2[j]);
    }
    stv(DY, h, w, pk);
  });
  if (threadIdx.x >= kConsumers) return;          // (consumers only from here: named barrier 1)
  // the ring is drained: reuse its first bytes as the [lanes][C][2] reduction scratch
  float* red = reinterpret_cast<float*>(ring);
  named_bar_sync(1, kConsumers);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[(ln * p.C + c + j) * 2 + 0] = s1[j];
    red[(ln * p.C + c + j) * 2 + 1] = s2[j];
  }
  named_bar_sync(1, kConsumers);
  for (int ch = threadIdx.x; ch < p.C; ch += kConsumers) {
    float a1 = 0.f, a2 = 0.f;
    for (int l = 0; l < lanes; ++l) {
      const float2 t = *reinterpret_cast<const float2*>(&red[(l * p.C + ch) * 2]);
      a1 += t.x;
      a2 += t.y;
    }
    const float* st = p.stats + (static_cast<long long>(n) * p.C + ch) * 2;
    const float mean = st[0] * p.inv_count;
    const float rstd = rsqrtf(fmaxf(st[1] * p.inv_count - mean * mean, 0.f) + p.eps);
    float* bs = p.bstats + (static_cast<long long>(n) * p.C + ch) * 2;
    atomicAdd(bs, a1);
    atomicAdd(bs + 1, rstd * (a2 - mean * a1));
  }
}

template <bool MX>
__global__ void __launch_bounds__(kStreamThreads, 2) in_bwd_apply_stream_kernel(const InBwdApplyParams p,
                                                                               int pix_per_block, uint32_t rcp_w) {
  extern __shared__ __align__(128) uint8_t ring[];
  constexpr int NT = 2, STAGES = 6;
  const int c8 = p.C >> 3;
  const int lanes = kConsumers / c8;
  const int tid = threadIdx.x < kConsumers ? threadIdx.x : 0;
  const int c = (tid % c8) * 8;
  const int ln = tid / c8;
  const int n = blockIdx.y;
  StreamGeom g;
  g.W = p.W; g.C = p.C; g.rcp_w = rcp_w; g.spx = lanes * kPxT;
  g.pix0 = blockIdx.x * pix_per_block;
  g.pix1 = min(p.H * p.W, g.pix0 + pix_per_block);
  V32 views[NT];
  views[0] = v32(p.dy, n, 0);
  views[1] = v32(p.x, n, 0);
  if (blockIdx.x == 0 && blockIdx.y == 0 && p.dgamma != nullptr && threadIdx.x < kConsumers) {
    for (int cc = threadIdx.x; cc < p.C; cc += kConsumers) {   // parameter gradients: sum over images
      float dg = 0.f, db = 0.f;
      for (int nn = 0; nn < p.N; ++nn) {
        const float* bs = p.bstats + (static_cast<long long>(nn) * p.C + cc) * 2;
        db += bs[0];
        dg += bs[1];
      }
      p.dgamma[cc] += dg;
      p.dbeta[cc] += db;
    }
  }
  float kA[8], kB[8], kC[8];
  if (threadIdx.x < kConsumers) {
    float mean[8], rstd[8], ga[8], b0[8], b1[8];
    mean_rstd8(p.stats + (static_cast<long long>(n) * p.C + c) * 2, p.inv_count, p.eps, mean, rstd);
    load_f8(p.gamma + c, ga);
    load_f8(p.bstats + (static_cast<long long>(n) * p.C + c) * 2, b0);
    load_f8(p.bstats + (static_cast<long long>(n) * p.C + c) * 2 + 8, b1);
    const float bs[16] = {b0[0], b0[1], b0[2], b0[3], b0[4], b0[5], b0[6], b0[7],
                          b1[0], b1[1], b1[2], b1[3], b1[4], b1[5], b1[6], b1[7]};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float m1 = bs[2 * j] * p.inv_count, m2 = bs[2 * j + 1] * p.inv_count;
      const float gr = ga[j] * rstd[j];
      kA[j] = gr;
      kB[j] = -gr * rstd[j] * m2;
      kC[j] = gr * (mean[j] * rstd[j] * m2 - m1);
    }
  }
  const V32 DX = v32(p.dx, n, c);
  uint8_t* d8 = nullptr;
  uint8_t* sf = nullptr;
  int d8sh = 0, d8sw = 0, sfsh = 0, sfsw = 0;
  if constexpr (MX) {
    d8 = reinterpret_cast<uint8_t*>(p.dx8.p) + static_cast<long long>(n) * p.dx8.sn + c;
    d8sh = static_cast<int>(p.dx8.sh);
    d8sw = static_cast<int>(p.dx8.sw);
    sf = reinterpret_cast<uint8_t*>(p.sf8.p) + static_cast<long long>(n) * p.sf8.sn;
    sfsh = static_cast<int>(p.sf8.sh);
    sfsw = static_cast<int>(p.sf8.sw);
  }
  const bool lane0 = (threadIdx.x & 31) == 0;
  stream_run<NT, STAGES>(ring, views, g, c, ln, lanes, [&](int, int h, int w, const uint4 (&v)[NT]) {
    float d[8], x[8], o[8];
    unpack8(v[0], d);
    unpack8(v[1], x);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(kA[j], d[j], fmaf(kB[j], x[j], kC[j]));
    stv(DX, h, w, pack8(o));
    if constexpr (MX) {   // C == 256: this warp == this pixel
      uint2 pk8, rec;
      mx_quant_pixel(o, pk8, rec);
      *reinterpret_cast<uint2*>(d8 + (h * d8sh + w * d8sw)) = pk8;
      if (lane0) *reinterpret_cast<uint2*>(sf + (h * sfsh + w * sfsw)) = rec;
    }
  });
}

// ------------------------------------------------------------------ act_bwd
__global__ void __launch_bounds__(kThreads) act_bwd_kernel(const ActBwdParams p) {
  __shared__ float sb[64];
  const uint32_t c8 = static_cast<uint32_t>(p.C >> 3);
  const uint32_t total = static_cast<uint32_t>(p.N) * p.H * p.W * c8;
  if (threadIdx.x < 64) sb[threadIdx.x] = 0.f;
  __syncthreads();
  float bacc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bacc[j] = 0.f;
  int my_c = -1;
  const bool is_tanh = p.act == ACT_TANH;
  const float slope = act_slope(p.act);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t pix = i / c8;
    const int c = static_cast<int>(i - pix * c8) * 8;
    int n, h, w;
    unravel(pix, p.W, p.H, n, h, w);
    float g[8], o[8];
    gather_g(p.g1, p.fold1, p.g2, n, h, w, p.H, p.W, c, g);
    load8(vptr(p.out, n, h, w) + c, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = is_tanh ? 1.f - o[j] * o[j] : (o[j] > 0.f ? 1.f : slope);
      g[j] = (c + j < p.c_valid) ? g[j] * d : 0.f;
      bacc[j] += __bfloat162float(__float2bfloat16(g[j]));
    }
    my_c = c;   // grid stride is a multiple of c8 => a thread always sees the same c
    store8(vptr_mut(p.dy, n, h, w) + c, g);
  }
  if (p.dbias != nullptr) {
    // lanes l and l + k * c8 of a warp hold the same channel group: combine them with shuffles
    // first (blockDim and the grid stride are multiples of c8), then one shared atomic per warp
    if (c8 <= 32) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        for (uint32_t o = 16; o >= c8; o >>= 1) bacc[j] += __shfl_xor_sync(0xffffffffu, bacc[j], o);
    }
    const int lane_c = static_cast<int>((threadIdx.x & 31u) % c8) * 8;
    if ((threadIdx.x & 31u) < c8 || c8 > 32) {
      const int cc = c8 <= 32 ? lane_c : my_c;
      if (cc >= 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (cc + j < p.c_valid) atomicAdd(&sb[cc + j], bacc[j]);
      }
    }
    __syncthreads();
    if (threadIdx.x < p.c_valid) atomicAdd(p.dbias + threadIdx.x, sb[threadIdx.x]);
  }
}

// ----------------------------------------------------------------- img_prep
__global__ void __launch_bounds__(kThreads) img_prep_kernel(const ImgPrepParams p) {
  const long long total = static_cast<long long>(p.N) * p.H * p.W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int n, h, w;
    unravel(static_cast<uint32_t>(i), p.W, p.H, n, h, w);
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (p.src_is_f32) {
      const float* s = reinterpret_cast<const float*>(p.src) + i * 3;
      v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
    } else {
      float t[8];
      load8(vptr(p.src_view, n, h, w), t);
      v[0] = t[0]; v[1] = t[1]; v[2] = t[2];
    }
    const uint4 pk = pack8(v);
    if (p.dst_plain.p != nullptr)
      *reinterpret_cast<uint4*>(vptr_mut(p.dst_plain, n, h, w)) = pk;
    if (p.dst_reflect.p != nullptr) {
      int hs[2], ws[2];
      const int nh = reflect_set(h, p.H, p.reflect_pad, hs);
      const int nw = reflect_set(w, p.W, p.reflect_pad, ws);
      for (int a = 0; a < nh; ++a)
        for (int b = 0; b < nw; ++b)
          *reinterpret_cast<uint4*>(vptr_mut(p.dst_reflect, n, hs[a], ws[b])) = pk;
    }
  }
}

// ------------------------------------------------------------------- losses
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float sh[kThreads / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < kThreads / 32) t = sh[threadIdx.x];
  if (threadIdx.x < 32) t = warp_sum(t);
  __syncthreads();
  return t;   // valid in thread 0
}

__global__ void __launch_bounds__(kThreads) l1_loss_kernel(const L1LossParams p) {
  const long long total = static_cast<long long>(p.N) * p.H * p.W;
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int n, h, w;
    unravel(static_cast<uint32_t>(i), p.W, p.H, n, h, w);
    float pr[8];
    load8(vptr(p.pred, n, h, w), pr);
    const float* t = p.target + i * 3;
    float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float d = t[j] - pr[j];
      acc += fabsf(d);
      g[j] = d > 0.f ? -p.grad_scale : (d < 0.f ? p.grad_scale : 0.f);
    }
    if (p.g.p != nullptr) store8(vptr_mut(p.g, n, h, w), g);
  }
  const float tot = block_sum(acc);
  if (threadIdx.x == 0) {
    atomicAdd(p.loss_slot, tot * p.loss_scale);
    if (p.err_slot != nullptr) atomicAdd(p.err_slot, tot * p.err_scale);
  }
}

__global__ void __launch_bounds__(kThreads) mse_loss_kernel(const MseLossParams p) {
  const long long total = static_cast<long long>(p.N) * p.H * p.W;
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int n, h, w;
    unravel(static_cast<uint32_t>(i), p.W, p.H, n, h, w);
    const float d = __bfloat162float(*vptr(p.d, n, h, w));
    const float e = p.target - d;
    acc += e * e;
    if (p.g.p != nullptr) *vptr_mut(p.g, n, h, w) = __float2bfloat16(-p.grad_scale * e);
  }
  const float tot = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(p.loss_slot, tot * p.loss_scale);
}

// --------------------------------------------------------------- tap gather
__global__ void __launch_bounds__(kThreads) tap_gather_kernel(const TapGatherParams p) {
  const long long total = static_cast<long long>(p.N) * p.H * p.W;
  const bool is_tanh = p.act == ACT_TANH;
  const float b0 = p.bias ? p.bias[0] : 0.f, b1 = p.bias ? p.bias[1] : 0.f, b2 = p.bias ? p.bias[2] : 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int n, h, w;
    unravel(static_cast<uint32_t>(i), p.W, p.H, n, h, w);
    const long long row0 = (static_cast<long long>(n) * p.Hp + h) * p.Wp + w;
    float a0 = b0, a1 = b1, a2 = b2;
#pragma unroll 7
    for (int s = 0; s < p.k; ++s) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(p.q + (row0 + s) * 32 + s * 4));
      const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
      a0 += lo.x; a1 += lo.y; a2 += hi.x;
    }
    float v[8] = {a0, a1, a2, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (is_tanh) { v[0] = fast_tanh(v[0]); v[1] = fast_tanh(v[1]); v[2] = fast_tanh(v[2]); }
    store8(vptr_mut(p.out, n, h, w), v);
  }
}

// --------------------------------------------------------------------- adam
__global__ void __launch_bounds__(kThreads) adam_kernel(const AdamParams p) {
  const int t = *p.step + 1;
  const float lr_t = p.lr * sqrtf(1.f - powf(p.beta2, static_cast<float>(t))) /
                     (1.f - powf(p.beta1, static_cast<float>(t)));
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < p.n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float g = p.g[i] * p.grad_scale;
    const float m = p.beta1 * p.m[i] + (1.f - p.beta1) * g;
    const float v = p.beta2 * p.v[i] + (1.f - p.beta2) * g * g;
    const float w = p.w[i] - lr_t * m / (sqrtf(v) + p.eps);
    p.m[i] = m;
    p.v[i] = v;
    p.w[i] = w;
    if (p.w_bf16 != nullptr) p.w_bf16[i] = __float2bfloat16(w);
  }
}
__global__ void step_increment_kernel(int* step) { *step += 1; }

__global__ void __launch_bounds__(kThreads) mask_grad_kernel(float* g, const uint8_t* mask,
                                                             long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    if (mask[i] == 0) g[i] = 0.f;
}

// ------------------------------------------------------------- FP8 bookkeeping
__global__ void __launch_bounds__(kThreads) fp8_w_amax_kernel(const Fp8Table t) {
  const int l = blockIdx.y;
  const __nv_bfloat16* w = t.w + t.w_off[l];
  float m = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < t.w_numel[l];
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    m = fmaxf(m, fabsf(__bfloat162float(w[i])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(t.w_amax + l), __float_as_int(m));
}
__device__ __forceinline__ float pow2_scale(float amax, float target) {
  return exp2f(floorf(log2f(target / amax)));
}
__global__ void fp8_scales_kernel(const Fp8Table t) {
  const int l = threadIdx.x;
  if (l >= t.num_layers) return;
  const float a = t.act_amax[l];
  const float sa = a > 0.f ? pow2_scale(a, 224.f) : t.default_act_scale;   // E4M3 max = 448: 2x headroom
  const float wa = t.w_amax[l];
  const float sb = wa > 0.f ? pow2_scale(wa, 224.f) : t.w_scale[l];
  t.act_scale[l] = sa;
  t.w_scale[l] = sb;
  t.out_scale[l] = 1.f / (sa * sb);
  t.act_amax[l] = 0.f;
}
__global__ void __launch_bounds__(kThreads) fp8_w_quant_kernel(const Fp8Table t) {
  const int l = blockIdx.y;
  const __nv_bfloat16* w = t.w + t.w_off[l];
  uint8_t* w8 = t.w8 + t.w_off[l];
  const float sb = t.w_scale[l];
  const long long n2 = t.w_numel[l] / 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n2;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(w + 2 * i));
    *reinterpret_cast<uint16_t*>(w8 + 2 * i) =
        __nv_cvt_float2_to_fp8x2(make_float2(v.x * sb, v.y * sb), __NV_SATFINITE, __NV_E4M3);
  }
}

// --------------------------------------------------- PatchGAN output layer bwd
// dX[n,h,w,c] = sum_{r,s} gd[n, h - r + pb, w - s + pb] * W[0][r*4+s][c]
__global__ void __launch_bounds__(kThreads) d4_dgrad_kernel(const D4BwdParams p) {
  const uint32_t c8 = static_cast<uint32_t>(p.C >> 3);
  const uint32_t total = static_cast<uint32_t>(p.N) * p.H * p.W * c8;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t pix = i / c8;
    const int c = static_cast<int>(i - pix * c8) * 8;
    int n, h, w;
    unravel(pix, p.W, p.H, n, h, w);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int kr = 0; kr < 4; ++kr) {
      const int oh = h - kr + p.pad_before;
      if (oh < 0 || oh >= p.H) continue;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int ow = w - ks + p.pad_before;
        if (ow < 0 || ow >= p.W) continue;
        const float g = __bfloat162float(*vptr(p.gd, n, oh, ow));
        float wv[8];
        load8(p.w + (kr * 4 + ks) * p.C + c, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += g * wv[j];
      }
    }
    store8(vptr_mut(p.dx, n, h, w) + c, acc);
  }
}

// dW[0][t][c] += sum_q gd[q - tap_t] * a[q][c]. The activation is read ONCE: a thread owns 8
// channels (one 16-byte load per pixel) and keeps the 16 tap accumulators in registers; the
// 256 / (C/8) pixel lanes of a block are combined in shared memory, then one atomic per value.
// grid: pixel chunks. dbias += sum gd.
__global__ void __launch_bounds__(kThreads) d4_wgrad_kernel(const D4BwdParams p, int pix_per_block) {
  extern __shared__ float red[];   // [16][C]
  const int c8 = p.C >> 3;
  const int lanes = kThreads / c8;
  const int c = (threadIdx.x % c8) * 8;
  const int ln = threadIdx.x / c8;
  const uint32_t npix = static_cast<uint32_t>(p.N) * p.H * p.W;
  const uint32_t q0 = blockIdx.x * static_cast<uint32_t>(pix_per_block);
  const uint32_t q1 = min(npix, q0 + static_cast<uint32_t>(pix_per_block));
  for (int i = threadIdx.x; i < 16 * p.C; i += kThreads) red[i] = 0.f;
  __syncthreads();
  float acc[16][8];
#pragma unroll
  for (int t = 0; t < 16; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  float gsum = 0.f;
  for (uint32_t q = q0 + ln; q < q1; q += lanes) {
    int n, h, w;
    unravel(q, p.W, p.H, n, h, w);
    float a[8];
    load8(vptr(p.a, n, h, w) + c, a);
    if (c == 0) gsum += __bfloat162float(*vptr(p.gd, n, h, w));
#pragma unroll
    for (int kr = 0; kr < 4; ++kr) {
      const int oh = h - kr + p.pad_before;      // output pixel that reads a[q] through tap (kr, ks)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int ow = w - ks + p.pad_before;
        float g = 0.f;
        if (oh >= 0 && oh < p.H && ow >= 0 && ow < p.W) g = __bfloat162float(*vptr(p.gd, n, oh, ow));
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[kr * 4 + ks][j] += g * a[j];
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 16; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&red[t * p.C + c + j], acc[t][j]);
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * p.C; i += kThreads) atomicAdd(p.dw + i, red[i]);
  if (p.dbias != nullptr && c == 0) atomicAdd(p.dbias, gsum);
}

inline int grid_for(long long total, int cap = 148 * 16) {
  long long b = (total + kThreads - 1) / kThreads;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}
// grid size that is a multiple of c8 channel groups so a thread keeps the same channel group
inline int grid_for_fixed_c(long long total, int c8) {
  int g = grid_for(total);
  // stride = g * 256 must be a multiple of c8 (c8 is a power of two <= 64 => always true)
  (void)c8;
  return g;
}

}  // namespace

static int ew_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Pixels per block of the (pixel chunks, N) grids. The kernels keep `blocks_per_sm` blocks
// resident, so the grid is sized to fill whole waves of the machine: every image gets
// floor(slots / N) chunks (chunk length need not be a power of two, only a multiple of the pixel
// lanes), e.g. 12 images x 49 chunks = 588 blocks on the 148 x 4 = 592 resident slots. Longer chunks
// also amortise the per-block coefficient set-up / reduction.
static int pick_pix_per_block(int HW, int N, int lanes, int blocks_per_sm) {
  const int slots = ew_num_sms() * blocks_per_sm;
  int chunks = slots / (N > 0 ? N : 1);
  if (chunks < 1) chunks = 1;
  int ppb = (HW + chunks - 1) / chunks;
  const int min_ppb = lanes * 4;             // >= 4 pixels per thread
  if (ppb < min_ppb) ppb = min_ppb;
  ppb = (ppb + lanes - 1) / lanes * lanes;
  return ppb > HW ? HW : ppb;
}

// reciprocal for pixel -> row: floor(pix / W) == umulhi(pix, ceil(2^32 / W)) for pix * W < 2^32
static uint32_t rcp_for(int H, int W, int C, const char* who) {
  if (W < 2 || C % 8 != 0 || kThreads % (C / 8) != 0 ||
      static_cast<unsigned long long>(H) * W * W >= (1ull << 32)) {
    fprintf(stderr, "%s: unsupported shape H=%d W=%d C=%d\n", who, H, W, C);
    abort();
  }
  return static_cast<uint32_t>(((1ull << 32) + W - 1) / W);
}

// ---- streamed variants: used whenever every streamed view has pixel stride == C (true for every
// buffer of the engine); the direct kernels remain for the FP8-twin output and odd strides.
static bool g_ew_stream = true;
void set_ew_stream(bool on) { g_ew_stream = on; }
bool ew_stream() { return g_ew_stream; }

static bool streamable(const View& v, int C) { return v.p == nullptr || v.sw == C; }

static int stream_ppb(int HW, int N, int C) {
  const int spx = (kConsumers / (C / 8)) * kPxT;
  int chunks = ew_num_sms() * 2 / (N > 0 ? N : 1);      // two resident blocks per SM, one wave
  if (chunks < 1) chunks = 1;
  int ppb = (HW + chunks - 1) / chunks;
  ppb = (ppb + spx - 1) / spx * spx;
  if (ppb < 2 * spx) ppb = 2 * spx;
  return ppb > HW ? HW : ppb;
}
template <class K>
static void stream_smem(K kernel, size_t bytes) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
}
constexpr size_t ring_bytes(int nt, int stages) { return static_cast<size_t>(stages) * nt * kStageBytes + 2 * stages * 8; }

template <int ACT, bool SKIP, bool HALO>
static void launch_in_apply_stream_t(const InApplyParams& p, dim3 grid, int ppb, uint32_t rcp, cudaStream_t s) {
  constexpr size_t smem = ring_bytes(SKIP ? 2 : 1, SKIP ? 6 : 8);
  if (p.sf8.p != nullptr) {
    static bool once_mx = (stream_smem(in_apply_stream_kernel<ACT, SKIP, HALO, true>, smem), true);
    (void)once_mx;
    in_apply_stream_kernel<ACT, SKIP, HALO, true><<<grid, kStreamThreads, smem, s>>>(p, ppb, rcp);
    return;
  }
  static bool once = (stream_smem(in_apply_stream_kernel<ACT, SKIP, HALO, false>, smem), true);
  (void)once;
  in_apply_stream_kernel<ACT, SKIP, HALO, false><<<grid, kStreamThreads, smem, s>>>(p, ppb, rcp);
}
template <int ACT>
static void launch_in_apply_stream_a(const InApplyParams& p, dim3 grid, int ppb, uint32_t rcp, cudaStream_t s) {
  const bool skip = p.skip.p != nullptr, halo = p.halo > 0;
  if (skip) { if (halo) launch_in_apply_stream_t<ACT, true, true>(p, grid, ppb, rcp, s); else launch_in_apply_stream_t<ACT, true, false>(p, grid, ppb, rcp, s); }
  else { if (halo) launch_in_apply_stream_t<ACT, false, true>(p, grid, ppb, rcp, s); else launch_in_apply_stream_t<ACT, false, false>(p, grid, ppb, rcp, s); }
}

template <int ACT, bool SKIP, bool HALO>
static void launch_in_apply_t(const InApplyParams& p, dim3 grid, int ppb, uint32_t rcp, cudaStream_t s) {
  if (p.dst8.p != nullptr && p.sf8.p != nullptr) in_apply_kernel<ACT, SKIP, HALO, 2><<<grid, kThreads, 0, s>>>(p, ppb, rcp);
  else if (p.dst8.p != nullptr) in_apply_kernel<ACT, SKIP, HALO, 1><<<grid, kThreads, 0, s>>>(p, ppb, rcp);
  else in_apply_kernel<ACT, SKIP, HALO, 0><<<grid, kThreads, 0, s>>>(p, ppb, rcp);
}
template <int ACT>
static void launch_in_apply_a(const InApplyParams& p, dim3 grid, int ppb, uint32_t rcp, cudaStream_t s) {
  const bool skip = p.skip.p != nullptr, halo = p.halo > 0;
  if (skip) { if (halo) launch_in_apply_t<ACT, true, true>(p, grid, ppb, rcp, s); else launch_in_apply_t<ACT, true, false>(p, grid, ppb, rcp, s); }
  else { if (halo) launch_in_apply_t<ACT, false, true>(p, grid, ppb, rcp, s); else launch_in_apply_t<ACT, false, false>(p, grid, ppb, rcp, s); }
}
void launch_in_apply(const InApplyParams& p, cudaStream_t s) {
  if (p.sf8.p != nullptr && (p.dst8.p == nullptr || p.C != 256)) {
    fprintf(stderr, "launch_in_apply: the MXFP8 twin needs dst8 and C == 256\n");
    abort();
  }
  const uint32_t rcp = rcp_for(p.H, p.W, p.C, "launch_in_apply");
  const int HW = p.H * p.W;
  if (g_ew_stream && (p.dst8.p == nullptr || p.sf8.p != nullptr) && streamable(p.x, p.C) && streamable(p.skip, p.C)) {
    const int ppb = stream_ppb(HW, p.N, p.C);
    dim3 grid((HW + ppb - 1) / ppb, p.N);
    if (p.act == ACT_RELU) launch_in_apply_stream_a<ACTK_RELU>(p, grid, ppb, rcp, s);
    else if (p.act == ACT_NONE) launch_in_apply_stream_a<ACTK_NONE>(p, grid, ppb, rcp, s);
    else launch_in_apply_stream_a<ACTK_SLOPE>(p, grid, ppb, rcp, s);
    return;
  }
  const int lanes = kThreads / (p.C / 8);
  const int ppb = pick_pix_per_block(HW, p.N, lanes, (p.skip.p != nullptr || p.dst8.p != nullptr) ? 3 : 4);
  dim3 grid((HW + ppb - 1) / ppb, p.N);
  if (p.act == ACT_RELU) launch_in_apply_a<ACTK_RELU>(p, grid, ppb, rcp, s);
  else if (p.act == ACT_NONE) launch_in_apply_a<ACTK_NONE>(p, grid, ppb, rcp, s);
  else launch_in_apply_a<ACTK_SLOPE>(p, grid, ppb, rcp, s);
}

template <int ACT, bool FOLD, bool G2>
static void launch_in_bwd_reduce_stream_t(const InBwdReduceParams& p, dim3 grid, int ppb, uint32_t rcp, cudaStream_t s) {
  constexpr size_t smem = ring_bytes(G2 ? 3 : 2, G2 ? 4 : 6);
  static bool once = (stream_smem(in_bwd_reduce_stream_kernel<ACT, FOLD, G2>, smem), true);
  (void)once;
  in_bwd_reduce_stream_kernel<ACT, FOLD, G2><<<grid, kStreamThreads, smem, s>>>(p, ppb, rcp);
}
template <int ACT>
static void launch_in_bwd_reduce_stream_a(const InBwdReduceParams& p, dim3 grid, int ppb, uint32_t rcp, cudaStream_t s) {
  const bool fold = p.fold1 > 0, g2 = p.g2.p != nullptr;
  if (fold) { if (g2) launch_in_bwd_reduce_stream_t<ACT, true, true>(p, grid, ppb, rcp, s); else launch_in_bwd_reduce_stream_t<ACT, true, false>(p, grid, ppb, rcp, s); }
  else { if (g2) launch_in_bwd_reduce_stream_t<ACT, false, true>(p, grid, ppb, rcp, s); else launch_in_bwd_reduce_stream_t<ACT, false, false>(p, grid, ppb, rcp, s); }
}
template <int ACT, bool FOLD, bool G2>
static void launch_in_bwd_reduce_t(const InBwdReduceParams& p, dim3 grid, size_t smem, int ppb, uint32_t rcp,
                                   cudaStream_t s) {
  if (p.stats != nullptr) in_bwd_reduce_kernel<ACT, FOLD, G2, true><<<grid, kThreads, smem, s>>>(p, ppb, rcp);
  else in_bwd_reduce_kernel<ACTK_NONE, FOLD, G2, false><<<grid, kThreads, smem, s>>>(p, ppb, rcp);
}
template <int ACT>
static void launch_in_bwd_reduce_a(const InBwdReduceParams& p, dim3 grid, size_t smem, int ppb, uint32_t rcp,
                                   cudaStream_t s) {
  const bool fold = p.fold1 > 0, g2 = p.g2.p != nullptr;
  if (fold) { if (g2) launch_in_bwd_reduce_t<ACT, true, true>(p, grid, smem, ppb, rcp, s); else launch_in_bwd_reduce_t<ACT, true, false>(p, grid, smem, ppb, rcp, s); }
  else { if (g2) launch_in_bwd_reduce_t<ACT, false, true>(p, grid, smem, ppb, rcp, s); else launch_in_bwd_reduce_t<ACT, false, false>(p, grid, smem, ppb, rcp, s); }
}
void launch_in_bwd_reduce(const InBwdReduceParams& p, cudaStream_t s) {
  const uint32_t rcp = rcp_for(p.H, p.W, p.C, "launch_in_bwd_reduce");
  const int HW = p.H * p.W;
  const int lanes = kThreads / (p.C / 8);
  if (g_ew_stream && p.stats != nullptr && p.bstats != nullptr && streamable(p.g1, p.C) && streamable(p.g2, p.C) &&
      streamable(p.x, p.C) && static_cast<size_t>(lanes) * p.C * 8 <= 4 * kStageBytes) {
    const int ppb = stream_ppb(HW, p.N, p.C);
    dim3 grid((HW + ppb - 1) / ppb, p.N);
    if (p.act == ACT_NONE) launch_in_bwd_reduce_stream_a<ACTK_NONE>(p, grid, ppb, rcp, s);
    else launch_in_bwd_reduce_stream_a<ACTK_SLOPE>(p, grid, ppb, rcp, s);
    return;
  }
  const int ppb = pick_pix_per_block(HW, p.N, lanes, 3);
  dim3 grid((HW + ppb - 1) / ppb, p.N);
  const size_t smem = static_cast<size_t>(lanes) * p.C * 2 * sizeof(float);
  if (p.act == ACT_NONE) launch_in_bwd_reduce_a<ACTK_NONE>(p, grid, smem, ppb, rcp, s);
  else launch_in_bwd_reduce_a<ACTK_SLOPE>(p, grid, smem, ppb, rcp, s);
}

void launch_in_bwd_apply(const InBwdApplyParams& p, cudaStream_t s) {
  const uint32_t rcp = rcp_for(p.H, p.W, p.C, "launch_in_bwd_apply");
  const int HW = p.H * p.W;
  const bool mx = p.sf8.p != nullptr;
  if (mx && (p.dx8.p == nullptr || p.C != 256 || !streamable(p.dy, p.C) || !streamable(p.x, p.C))) {
    fprintf(stderr, "launch_in_bwd_apply: the MXFP8 twin needs dx8, C == 256 and dense dy / x views\n");
    abort();
  }
  if ((g_ew_stream || mx) && streamable(p.dy, p.C) && streamable(p.x, p.C)) {
    constexpr size_t smem = ring_bytes(2, 6);
    static bool once = (stream_smem(in_bwd_apply_stream_kernel<false>, smem),
                        stream_smem(in_bwd_apply_stream_kernel<true>, smem), true);
    (void)once;
    const int ppb = stream_ppb(HW, p.N, p.C);
    dim3 grid((HW + ppb - 1) / ppb, p.N);
    if (mx) in_bwd_apply_stream_kernel<true><<<grid, kStreamThreads, smem, s>>>(p, ppb, rcp);
    else in_bwd_apply_stream_kernel<false><<<grid, kStreamThreads, smem, s>>>(p, ppb, rcp);
    return;
  }
  const int lanes = kThreads / (p.C / 8);
  const int ppb = pick_pix_per_block(HW, p.N, lanes, 3);
  dim3 grid((HW + ppb - 1) / ppb, p.N);
  in_bwd_apply_kernel<<<grid, kThreads, 0, s>>>(p, ppb, rcp);
}

void launch_act_bwd(const ActBwdParams& p, cudaStream_t s) {
  const long long total = static_cast<long long>(p.N) * p.H * p.W * (p.C / 8);
  // every block ends with c_valid same-address global atomics (dbias): keep the grid small
  act_bwd_kernel<<<grid_for(total, ew_num_sms() * 2), kThreads, 0, s>>>(p);
}

void launch_img_prep(const ImgPrepParams& p, cudaStream_t s) {
  const long long total = static_cast<long long>(p.N) * p.H * p.W;
  img_prep_kernel<<<grid_for(total), kThreads, 0, s>>>(p);
}

void launch_l1_loss(const L1LossParams& p, cudaStream_t s) {
  const long long total = static_cast<long long>(p.N) * p.H * p.W;
  l1_loss_kernel<<<grid_for(total, 148 * 4), kThreads, 0, s>>>(p);
}

void launch_mse_loss(const MseLossParams& p, cudaStream_t s) {
  const long long total = static_cast<long long>(p.N) * p.H * p.W;
  mse_loss_kernel<<<grid_for(total, 148), kThreads, 0, s>>>(p);
}

void launch_fp8_w_amax(const Fp8Table& t, cudaStream_t s) {
  fp8_w_amax_kernel<<<dim3(64, t.num_layers), kThreads, 0, s>>>(t);
}
void launch_fp8_scales(const Fp8Table& t, cudaStream_t s) { fp8_scales_kernel<<<1, 32, 0, s>>>(t); }
void launch_fp8_w_quant(const Fp8Table& t, cudaStream_t s) {
  fp8_w_quant_kernel<<<dim3(64, t.num_layers), kThreads, 0, s>>>(t);
}

void launch_tap_gather(const TapGatherParams& p, cudaStream_t s) {
  const long long total = static_cast<long long>(p.N) * p.H * p.W;
  tap_gather_kernel<<<grid_for(total), kThreads, 0, s>>>(p);
}

void launch_adam(const AdamParams& p, cudaStream_t s) {
  adam_kernel<<<grid_for(p.n, 148 * 8), kThreads, 0, s>>>(p);
}
void launch_step_increment(int* step, cudaStream_t s) { step_increment_kernel<<<1, 1, 0, s>>>(step); }

void launch_mask_grad(float* g, const uint8_t* mask, long long n, cudaStream_t s) {
  mask_grad_kernel<<<grid_for(n), kThreads, 0, s>>>(g, mask, n);
}

void launch_d4_dgrad(const D4BwdParams& p, cudaStream_t s) {
  const long long total = static_cast<long long>(p.N) * p.H * p.W * (p.C / 8);
  d4_dgrad_kernel<<<grid_for(total), kThreads, 0, s>>>(p);
}
void launch_d4_wgrad(const D4BwdParams& p, cudaStream_t s) {
  const int npix = p.N * p.H * p.W;
  const int lanes = kThreads / (p.C / 8);
  int ppb = (npix + ew_num_sms() - 1) / ew_num_sms();      // one block per SM
  ppb = (ppb + lanes - 1) / lanes * lanes;
  const size_t smem = static_cast<size_t>(16) * p.C * sizeof(float);
  if (p.C % 8 != 0 || kThreads % (p.C / 8) != 0 || smem > 48 * 1024) {
    fprintf(stderr, "launch_d4_wgrad: unsupported channel count %d\n", p.C);
    abort();
  }
  d4_wgrad_kernel<<<(npix + ppb - 1) / ppb, kThreads, smem, s>>>(p, ppb);
}

}  // namespace b2
