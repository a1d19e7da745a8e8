// Bandwidth-bound kernels around the GEMMs (see elementwise.cuh).
#include <stdlib.h>

#include <cuda_fp8.h>

#include "elementwise.cuh"

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

// ------------------------------------------------------------------ in_apply
// grid: (pixel chunks, N). A thread owns one 8-channel group (coefficients live in
// registers) and walks pixels, 4 at a time so that several 16-byte loads are in flight.
// U = pixels in flight per thread, MINB = resident blocks per SM the register budget is set for
// (variants are selected at run time, see ew_variant(); <4, 4> is the default)
template <int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) in_apply_kernel(const InApplyParams p, int pix_per_block) {
  const int c8 = p.C >> 3;
  const int lanes = kThreads / c8;
  const int c = (threadIdx.x % c8) * 8;
  const int ln = threadIdx.x / c8;
  const int n = blockIdx.y;
  const int HW = p.H * p.W;
  const int pix0 = blockIdx.x * pix_per_block;
  const int pix1 = min(HW, pix0 + pix_per_block);
  // per-channel affine coefficients (a, b), computed once per block (thread t -> channel t) and
  // read back from shared memory at use: few registers -> 4 blocks / SM -> more loads in flight
  extern __shared__ float coef[];   // float2[C]
  float2* ab = reinterpret_cast<float2*>(coef);
  for (int ch = threadIdx.x; ch < p.C; ch += kThreads) {
    float a = 1.f, b = 0.f;
    if (p.stats != nullptr) {
      const float* st = p.stats + (static_cast<long long>(n) * p.C + ch) * 2;
      const float mean = st[0] * p.inv_count;
      const float var = fmaxf(st[1] * p.inv_count - mean * mean, 0.f);
      a = rsqrtf(var + p.eps) * __ldg(p.gamma + ch);
      b = __ldg(p.beta + ch) - mean * a;
    }
    ab[((ch & 7) >> 1) * 2 * c8 + (ch >> 3) * 2 + (ch & 1)] = make_float2(a, b);
  }
  __syncthreads();
  const bool has_skip = p.skip.p != nullptr;
  const float slope = act_slope(p.act);
  const bool has8 = p.dst8.p != nullptr;
  const float s8 = has8 ? __ldg(p.scale8) : 1.f;
  float amax = 0.f;
  int hh[U], ww[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {   // one division per slot, then incremental (h, w) updates
    const int pix = pix0 + ln + u * lanes;
    hh[u] = pix / p.W;
    ww[u] = pix - hh[u] * p.W;
  }
  const int step = lanes * U;
  const int step_h = step / p.W, step_w = step - step_h * p.W;
  for (int base = pix0 + ln; base < pix1; base += step) {
    uint4 xv[U], sv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = base + u * lanes;
      if (pix < pix1) {
        xv[u] = __ldg(reinterpret_cast<const uint4*>(vptr(p.x, n, hh[u], ww[u]) + c));
        if (has_skip) sv[u] = __ldg(reinterpret_cast<const uint4*>(vptr(p.skip, n, hh[u], ww[u]) + c));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = base + u * lanes;
      if (pix >= pix1) continue;
      const uint32_t xw[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
      const uint32_t sw[4] = {sv[u].x, sv[u].y, sv[u].z, sv[u].w};
      uint32_t ow[4];
      uint16_t o8[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_bf16x2(xw[j]);
        const float4 k = *reinterpret_cast<const float4*>(&ab[(j * c8 + (c >> 3)) * 2]);   // (a0, b0, a1, b1)
        float v0 = apply_slope(fmaf(t.x, k.x, k.y), slope);
        float v1 = apply_slope(fmaf(t.y, k.z, k.w), slope);
        if (has_skip) {
          const float2 q = unpack_bf16x2(sw[j]);
          v0 += q.x;
          v1 += q.y;
        }
        ow[j] = pack_bf16x2(v0, v1);
        if (has8) {
          amax = fmaxf(amax, fmaxf(fabsf(v0), fabsf(v1)));
          o8[j] = __nv_cvt_float2_to_fp8x2(make_float2(v0 * s8, v1 * s8), __NV_SATFINITE, __NV_E4M3);
        }
      }
      const uint4 pk = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      *reinterpret_cast<uint4*>(vptr_mut(p.dst, n, hh[u], ww[u]) + c) = pk;
      uint2 pk8 = make_uint2(0u, 0u);
      if (has8) {
        pk8 = make_uint2(static_cast<uint32_t>(o8[0]) | (static_cast<uint32_t>(o8[1]) << 16),
                         static_cast<uint32_t>(o8[2]) | (static_cast<uint32_t>(o8[3]) << 16));
        *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(p.dst8.p) + n * p.dst8.sn + hh[u] * p.dst8.sh +
                                  ww[u] * p.dst8.sw + c) = pk8;
      }
      if (p.halo > 0 && (hh[u] <= p.halo || hh[u] >= p.H - 1 - p.halo || ww[u] <= p.halo ||
                         ww[u] >= p.W - 1 - p.halo)) {   // only border pixels have mirror images
        int hs[2], ws[2];
        const int nh = reflect_set(hh[u], p.H, p.halo, hs);
        const int nw = reflect_set(ww[u], p.W, p.halo, ws);
        if (nh + nw > 2)
          for (int a = 0; a < nh; ++a)
            for (int b = 0; b < nw; ++b)
              if (a + b > 0) {
                *reinterpret_cast<uint4*>(vptr_mut(p.dst, n, hs[a], ws[b]) + c) = pk;
                if (has8)
                  *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(p.dst8.p) + n * p.dst8.sn +
                                            hs[a] * p.dst8.sh + ws[b] * p.dst8.sw + c) = pk8;
              }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      hh[u] += step_h;
      ww[u] += step_w;
      if (ww[u] >= p.W) { ww[u] -= p.W; ++hh[u]; }
    }
  }
  if (has8) {   // delayed-scaling statistics: one atomic per warp (values are >= 0, so int order == float order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(p.amax8), __float_as_int(amax));
  }
}

// ------------------------------------------------------------ in_bwd_reduce
// grid: (pixel chunks, N). Threads: (C/8 channel groups) x (256 / (C/8) pixel lanes).
template <int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) in_bwd_reduce_kernel(const InBwdReduceParams p,
                                                                    int pix_per_block) {
  extern __shared__ float red[];   // [lanes][C][2]
  const int c8 = p.C >> 3;
  const int lanes = kThreads / c8;
  const int cg = threadIdx.x % c8;
  const int ln = threadIdx.x / c8;
  const int c = cg * 8;
  const int n = blockIdx.y;
  const int HW = p.H * p.W;
  const int pix0 = blockIdx.x * pix_per_block;
  const int pix1 = min(HW, pix0 + pix_per_block);

  const bool has_in = p.stats != nullptr;
  float4* mrgb = reinterpret_cast<float4*>(red + lanes * p.C * 2);   // float4[C] = (mean, rstd, gamma, beta)
  if (has_in) {
    for (int ch = threadIdx.x; ch < p.C; ch += kThreads) {
      const float* st = p.stats + (static_cast<long long>(n) * p.C + ch) * 2;
      const float mu = st[0] * p.inv_count;
      const float var = fmaxf(st[1] * p.inv_count - mu * mu, 0.f);
      mrgb[(ch & 7) * c8 + (ch >> 3)] = make_float4(mu, rsqrtf(var + p.eps), __ldg(p.gamma + ch), __ldg(p.beta + ch));
    }
    __syncthreads();
  }
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }

  const bool has_g2 = p.g2.p != nullptr;
  const float slope = act_slope(p.act);
  int hh[U], ww[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int pix = pix0 + ln + u * lanes;
    hh[u] = pix / p.W;
    ww[u] = pix - hh[u] * p.W;
  }
  const int step = lanes * U;
  const int step_h = step / p.W, step_w = step - step_h * p.W;
  for (int base = pix0 + ln; base < pix1; base += step) {
    uint4 gv[U], g2v[U], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = base + u * lanes;
      if (pix < pix1) {
        gv[u] = __ldg(reinterpret_cast<const uint4*>(vptr(p.g1, n, hh[u], ww[u]) + c));
        if (has_g2) g2v[u] = __ldg(reinterpret_cast<const uint4*>(vptr(p.g2, n, hh[u], ww[u]) + c));
        if (has_in) xv[u] = __ldg(reinterpret_cast<const uint4*>(vptr(p.x, n, hh[u], ww[u]) + c));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = base + u * lanes;
      if (pix >= pix1) continue;
      const int h = hh[u], w = ww[u];
      float g[8];
      {
        const uint32_t a[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
        const uint32_t b[4] = {g2v[u].x, g2v[u].y, g2v[u].z, g2v[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = unpack_bf16x2(a[j]);
          g[2 * j] = t.x; g[2 * j + 1] = t.y;
          if (has_g2) { const float2 k = unpack_bf16x2(b[j]); g[2 * j] += k.x; g[2 * j + 1] += k.y; }
        }
      }
      if (p.fold1 > 0 && (h <= p.fold1 || h >= p.H - 1 - p.fold1 || w <= p.fold1 ||
                          w >= p.W - 1 - p.fold1)) {   // rare: pixels whose mirror images live in the halo of g1
        int hs[2], ws[2];
        const int nh = reflect_set(h, p.H, p.fold1, hs);
        const int nw = reflect_set(w, p.W, p.fold1, ws);
        if (nh + nw > 2)
          for (int a = 0; a < nh; ++a)
            for (int b = 0; b < nw; ++b)
              if (a + b > 0) {
                float t[8];
                load8(vptr(p.g1, n, hs[a], ws[b]) + c, t);
#pragma unroll
                for (int j = 0; j < 8; ++j) g[j] += t[j];
              }
      }
      if (has_in) {
        const uint32_t xw[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = unpack_bf16x2(xw[j]);
          const float xs[2] = {t.x, t.y};
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int jj = 2 * j + q;
            const float4 k = mrgb[jj * c8 + (c >> 3)];
            const float xh = (xs[q] - k.x) * k.y;
            const float y = xh * k.z + k.w;
            g[jj] *= act_grad_from_pre(y, slope);
            const float gq = __bfloat162float(__float2bfloat16(g[jj]));   // stats of the stored value
            s1[jj] += gq;
            s2[jj] += gq * xh;
          }
        }
      }
      store8(vptr_mut(p.dy, n, h, w) + c, g);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      hh[u] += step_h;
      ww[u] += step_w;
      if (ww[u] >= p.W) { ww[u] -= p.W; ++hh[u]; }
    }
  }
  if (!has_in || p.bstats == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[(ln * p.C + c + j) * 2 + 0] = s1[j];
    red[(ln * p.C + c + j) * 2 + 1] = s2[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.C * 2; i += kThreads) {
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += red[l * p.C * 2 + i];
    atomicAdd(p.bstats + static_cast<long long>(n) * p.C * 2 + i, acc);
  }
}

// ------------------------------------------------------------- in_bwd_apply
// dx = A*dy + B*x + C per channel (A = gamma*rstd, B = -gamma*rstd^2*m2, C = gamma*rstd*(mean*rstd*m2 - m1))
template <int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) in_bwd_apply_kernel(const InBwdApplyParams p, int pix_per_block) {
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
  extern __shared__ float coef[];   // float4[C] = (A, B, C, -)
  float4* abc = reinterpret_cast<float4*>(coef);
  for (int ch = threadIdx.x; ch < p.C; ch += kThreads) {
    const float* st = p.stats + (static_cast<long long>(n) * p.C + ch) * 2;
    const float* bs = p.bstats + (static_cast<long long>(n) * p.C + ch) * 2;
    const float mean = st[0] * p.inv_count;
    const float var = fmaxf(st[1] * p.inv_count - mean * mean, 0.f);
    const float rstd = rsqrtf(var + p.eps);
    const float m1 = bs[0] * p.inv_count;
    const float m2 = bs[1] * p.inv_count;
    const float gr = __ldg(p.gamma + ch) * rstd;
    abc[(ch & 7) * c8 + (ch >> 3)] = make_float4(gr, -gr * rstd * m2, gr * (mean * rstd * m2 - m1), 0.f);
  }
  __syncthreads();
  int hh[U], ww[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int pix = pix0 + ln + u * lanes;
    hh[u] = pix / p.W;
    ww[u] = pix - hh[u] * p.W;
  }
  const int step = lanes * U;
  const int step_h = step / p.W, step_w = step - step_h * p.W;
  for (int base = pix0 + ln; base < pix1; base += step) {
    uint4 dv[U], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = base + u * lanes;
      if (pix < pix1) {
        dv[u] = __ldg(reinterpret_cast<const uint4*>(vptr(p.dy, n, hh[u], ww[u]) + c));
        xv[u] = __ldg(reinterpret_cast<const uint4*>(vptr(p.x, n, hh[u], ww[u]) + c));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = base + u * lanes;
      if (pix >= pix1) continue;
      const uint32_t dw[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
      const uint32_t xw[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 d = unpack_bf16x2(dw[j]);
        const float2 x = unpack_bf16x2(xw[j]);
        const float4 k0 = abc[(2 * j) * c8 + (c >> 3)], k1 = abc[(2 * j + 1) * c8 + (c >> 3)];
        ow[j] = pack_bf16x2(fmaf(k0.x, d.x, fmaf(k0.y, x.x, k0.z)), fmaf(k1.x, d.y, fmaf(k1.y, x.y, k1.z)));
      }
      *reinterpret_cast<uint4*>(vptr_mut(p.dx, n, hh[u], ww[u]) + c) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      hh[u] += step_h;
      ww[u] += step_w;
      if (ww[u] >= p.W) { ww[u] -= p.W; ++hh[u]; }
    }
  }
}

// Register double-buffered variant of in_bwd_apply (experiment, ew_variant() == 3): the loads of the
// next U pixels are issued before the current U are consumed, so a thread always has loads in
// flight while it computes. Same arithmetic as in_bwd_apply_kernel.
template <int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) in_bwd_apply_pf_kernel(const InBwdApplyParams p, int pix_per_block) {
  const int c8 = p.C >> 3;
  const int lanes = kThreads / c8;
  const int c = (threadIdx.x % c8) * 8;
  const int ln = threadIdx.x / c8;
  const int n = blockIdx.y;
  const int HW = p.H * p.W;
  const int pix0 = blockIdx.x * pix_per_block;
  const int pix1 = min(HW, pix0 + pix_per_block);
  if (blockIdx.x == 0 && blockIdx.y == 0 && p.dgamma != nullptr) {
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
  extern __shared__ float coef[];   // float4[C] = (A, B, C, -)
  float4* abc = reinterpret_cast<float4*>(coef);
  for (int ch = threadIdx.x; ch < p.C; ch += kThreads) {
    const float* st = p.stats + (static_cast<long long>(n) * p.C + ch) * 2;
    const float* bs = p.bstats + (static_cast<long long>(n) * p.C + ch) * 2;
    const float mean = st[0] * p.inv_count;
    const float var = fmaxf(st[1] * p.inv_count - mean * mean, 0.f);
    const float rstd = rsqrtf(var + p.eps);
    const float m1 = bs[0] * p.inv_count;
    const float m2 = bs[1] * p.inv_count;
    const float gr = __ldg(p.gamma + ch) * rstd;
    abc[(ch & 7) * c8 + (ch >> 3)] = make_float4(gr, -gr * rstd * m2, gr * (mean * rstd * m2 - m1), 0.f);
  }
  __syncthreads();
  const int step = lanes * U;
  const int step2 = 2 * step;
  const int s2h = step2 / p.W, s2w = step2 - s2h * p.W;
  int hA[U], wA[U], hB[U], wB[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int pa = pix0 + ln + u * lanes, pb = pa + step;
    hA[u] = pa / p.W; wA[u] = pa - hA[u] * p.W;
    hB[u] = pb / p.W; wB[u] = pb - hB[u] * p.W;
  }
  uint4 dA[U], xA[U], dB[U], xB[U];
  auto load = [&](uint4 (&dv)[U], uint4 (&xv)[U], const int (&hh)[U], const int (&ww)[U], int base) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (base + u * lanes < pix1) {
        dv[u] = __ldg(reinterpret_cast<const uint4*>(vptr(p.dy, n, hh[u], ww[u]) + c));
        xv[u] = __ldg(reinterpret_cast<const uint4*>(vptr(p.x, n, hh[u], ww[u]) + c));
      }
  };
  auto consume = [&](const uint4 (&dv)[U], const uint4 (&xv)[U], int (&hh)[U], int (&ww)[U], int base) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * lanes < pix1) {
        const uint32_t dw[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
        const uint32_t xw[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 d = unpack_bf16x2(dw[j]);
          const float2 x = unpack_bf16x2(xw[j]);
          const float4 k0 = abc[(2 * j) * c8 + (c >> 3)], k1 = abc[(2 * j + 1) * c8 + (c >> 3)];
          ow[j] = pack_bf16x2(fmaf(k0.x, d.x, fmaf(k0.y, x.x, k0.z)), fmaf(k1.x, d.y, fmaf(k1.y, x.y, k1.z)));
        }
        *reinterpret_cast<uint4*>(vptr_mut(p.dx, n, hh[u], ww[u]) + c) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
      hh[u] += s2h;                       // this buffer is next used 2 * step pixels further on
      ww[u] += s2w;
      if (ww[u] >= p.W) { ww[u] -= p.W; ++hh[u]; }
    }
  };
  int base = pix0 + ln;
  load(dA, xA, hA, wA, base);
  for (; base < pix1; base += step2) {
    load(dB, xB, hB, wB, base + step);
    consume(dA, xA, hA, wA, base);
    load(dA, xA, hA, wA, base + step2);
    consume(dB, xB, hB, wB, base + step);
  }
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

static int ew_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

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
// resident, so the grid is sized to fill whole waves of the machine: with B2_EW_WAVES = w
// (default 1) every image gets floor(w * slots / N) chunks (chunk length need not be a power of
// two, only a multiple of the pixel lanes), e.g. 12 images x 37 chunks = 444 blocks = exactly one
// wave of 148 x 3. Longer chunks also amortise the per-block coefficient table / reduction.
static int pick_pix_per_block(int HW, int N, int lanes, int blocks_per_sm) {
  static const int legacy = ew_env("B2_EW_BLOCKS", 0);
  if (legacy > 0) {   // old heuristic: power-of-two chunks until >= B2_EW_BLOCKS blocks
    int ppb = lanes * 32;
    while (ppb > lanes * 4 && static_cast<long long>((HW + ppb - 1) / ppb) * N < legacy) ppb /= 2;
    return ppb > HW ? HW : ppb;
  }
  static const int waves = ew_env("B2_EW_WAVES", 1);
  const int slots = ew_num_sms() * blocks_per_sm * (waves > 0 ? waves : 1);
  int chunks = slots / (N > 0 ? N : 1);
  if (chunks < 1) chunks = 1;
  int ppb = (HW + chunks - 1) / chunks;
  const int min_ppb = lanes * 4;             // >= 4 pixels per thread
  if (ppb < min_ppb) ppb = min_ppb;
  ppb = (ppb + lanes - 1) / lanes * lanes;
  return ppb > HW ? HW : ppb;
}

// Tuning variants of the three InstanceNorm kernels (pixels in flight per thread / resident blocks
// per SM / register double buffering); 0 = default. Same arithmetic in every variant.
static int g_ew_variant = -1;
int ew_variant() {
  if (g_ew_variant < 0) g_ew_variant = ew_env("B2_EW_VARIANT", 0);
  return g_ew_variant;
}
void set_ew_variant(int v) { g_ew_variant = v < 0 ? 0 : v; }

template <int U, int MINB>
static void launch_in_apply_t(const InApplyParams& p, cudaStream_t s) {
  const int HW = p.H * p.W;
  const int lanes = kThreads / (p.C / 8);
  const int ppb = pick_pix_per_block(HW, p.N, lanes, MINB);
  dim3 grid((HW + ppb - 1) / ppb, p.N);
  in_apply_kernel<U, MINB><<<grid, kThreads, 2 * p.C * sizeof(float), s>>>(p, ppb);
}
void launch_in_apply(const InApplyParams& p, cudaStream_t s) {
  switch (ew_variant()) {
    case 1: return launch_in_apply_t<2, 6>(p, s);
    case 2: return launch_in_apply_t<8, 3>(p, s);
    case 3: return launch_in_apply_t<4, 3>(p, s);
    default: return launch_in_apply_t<4, 4>(p, s);
  }
}

template <int U, int MINB>
static void launch_in_bwd_reduce_t(const InBwdReduceParams& p, cudaStream_t s) {
  const int HW = p.H * p.W;
  const int c8 = p.C / 8;
  const int lanes = kThreads / c8;
  const int pix_per_block = pick_pix_per_block(HW, p.N, lanes, MINB);
  dim3 grid((HW + pix_per_block - 1) / pix_per_block, p.N);
  const size_t smem = (static_cast<size_t>(lanes) * p.C * 2 + 4 * p.C) * sizeof(float);
  in_bwd_reduce_kernel<U, MINB><<<grid, kThreads, smem, s>>>(p, pix_per_block);
}
void launch_in_bwd_reduce(const InBwdReduceParams& p, cudaStream_t s) {
  switch (ew_variant()) {
    case 1: return launch_in_bwd_reduce_t<4, 2>(p, s);
    case 2: return launch_in_bwd_reduce_t<2, 4>(p, s);
    case 3: return launch_in_bwd_reduce_t<2, 2>(p, s);
    default: return launch_in_bwd_reduce_t<2, 3>(p, s);
  }
}

template <int U, int MINB, bool PF>
static void launch_in_bwd_apply_t(const InBwdApplyParams& p, cudaStream_t s) {
  const int HW = p.H * p.W;
  const int lanes = kThreads / (p.C / 8);
  const int ppb = pick_pix_per_block(HW, p.N, lanes, MINB);
  dim3 grid((HW + ppb - 1) / ppb, p.N);
  if constexpr (PF) in_bwd_apply_pf_kernel<U, MINB><<<grid, kThreads, 4 * p.C * sizeof(float), s>>>(p, ppb);
  else in_bwd_apply_kernel<U, MINB><<<grid, kThreads, 4 * p.C * sizeof(float), s>>>(p, ppb);
}
void launch_in_bwd_apply(const InBwdApplyParams& p, cudaStream_t s) {
  switch (ew_variant()) {
    case 1: return launch_in_bwd_apply_t<2, 6, false>(p, s);
    case 2: return launch_in_bwd_apply_t<8, 3, false>(p, s);
    case 3: return launch_in_bwd_apply_t<2, 4, true>(p, s);
    default: return launch_in_bwd_apply_t<4, 4, false>(p, s);
  }
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
