// Shared epilogue of the implicit-GEMM conv kernels: TMEM accumulator (128 lanes x BLOCK_N
// fp32 columns) -> bias / activation -> bf16 -> swizzled smem staging -> InstanceNorm partial
// statistics + coalesced global store, in 64-column chunks.
#pragma once
#include "conv_gemm.cuh"

namespace b2 {

template <int BLOCK_N>
struct EpiCfg {
  static constexpr int kChunkN = BLOCK_N < 64 ? BLOCK_N : 64;   // epilogue column chunk
  static constexpr int kChunks16 = kChunkN / 8;    // 16-byte pieces per staged row
  static constexpr int kRowGroups = 128 / kChunks16;
  static constexpr int kStagingBytes = 128 * kChunkN * 2;
  static constexpr int kRedBytes = kRowGroups * kChunkN * 2 * 4;
};

// Called by the 128 threads of ONE epilogue warpgroup (e = TMEM lane = tile row) for BLOCK_N
// accumulator columns starting at `taddr` / output channel n0; kernels run two warpgroups per
// tile, each on its own half of the columns with its own staging / red buffers and named
// barriers (bar_id, bar_id + 1). `release` is invoked once by every thread right after its last
// TMEM read (arrive on the tmem-empty barrier).
template <int BLOCK_N, class Release>
__device__ __forceinline__ void conv_epilogue_tile(const ConvGemmParams& p, __nv_bfloat16* out_base, uint8_t* stage_buf,
                                                   float* red, uint32_t taddr, int img, int h0,
                                                   int w0, int n0, int e, int ew, int lane,
                                                   bool tile_valid, Release release, int dbg = 0,
                                                   int bar_id = 1) {
  constexpr int CN = EpiCfg<BLOCK_N>::kChunkN;
  constexpr int CH = EpiCfg<BLOCK_N>::kChunks16;
  constexpr int RG = EpiCfg<BLOCK_N>::kRowGroups;
  constexpr bool SWZ = (CH >= 8);
  constexpr int ROW_BYTES = CN * 2;
  const bool has_bias = p.bias != nullptr;
  const bool is_tanh = p.act == ACT_TANH;
  const float slope = act_slope(p.act);
  const float oscale = p.out_scale != nullptr ? __ldg(p.out_scale) : 1.f;
  // global offsets of the rows this thread stores (same rows for every column chunk): hoists the
  // integer div / mod and 64-bit address math out of the chunk loop
  constexpr int RPI = 32 / (CH > 32 ? 32 : CH);   // rows per warp instruction
  constexpr int kStoreIters = 32 / RPI;
  long long row_off[kStoreIters];
#pragma unroll
  for (int itr = 0; itr < kStoreIters; ++itr) {
    const int r = ew * 32 + itr * RPI + lane / CH;
    const int wl = r % p.bw, hl = r / p.bw;
    row_off[itr] = (w0 + wl < p.w_limit && h0 + hl < p.h_limit)
                       ? img * p.out_sn + (h0 + hl) * p.out_sh + (w0 + wl) * p.out_sw
                       : -1;
  }
#pragma unroll 1
  for (int c0 = 0; c0 < BLOCK_N; c0 += CN) {
    // ---- TMEM -> registers -> bias / activation -> bf16 -> staging (swizzled rows)
    if constexpr (CN >= 32) {
#pragma unroll
      for (int cc = 0; cc < CN; cc += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c0 + cc, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * oscale;
        if (has_bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + cc + j < p.cout) v[j] += __ldg(p.bias + n0 + c0 + cc + j);
        }
        if (is_tanh) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fast_tanh(v[j]);
        } else if (slope != 1.f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = apply_slope(v[j], slope);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ci = (cc >> 3) + q;
          const int phys = SWZ ? (ci ^ (e & 7)) : ci;
          uint4 pk;
          pk.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
          pk.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
          pk.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
          pk.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
          *reinterpret_cast<uint4*>(stage_buf + e * ROW_BYTES + phys * 16) = pk;
        }
      }
    } else {
      uint32_t r[16];
      tmem_ld_32x16(taddr + c0, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) * oscale;
      if (has_bias) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n0 + c0 + j < p.cout) v[j] += __ldg(p.bias + n0 + c0 + j);
      }
      if (is_tanh) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fast_tanh(v[j]);
      } else if (slope != 1.f) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = apply_slope(v[j], slope);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint4 pk;
        pk.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
        pk.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
        pk.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
        pk.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
        *reinterpret_cast<uint4*>(stage_buf + e * ROW_BYTES + q * 16) = pk;
      }
    }
    if (c0 + CN >= BLOCK_N) {     // all TMEM reads of this tile are done: release the accumulator
      tc_fence_before();
      release();
    }
    named_bar_sync(bar_id, 128);

    // ---- InstanceNorm partial statistics (sum, sum of squares per channel)
    if (p.stats != nullptr && tile_valid) {
      const int ci = e % CH;
      const int rg = e / CH;
      float s[8], q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
#pragma unroll 4
      for (int r = rg; r < 128; r += RG) {
        const int phys = SWZ ? (ci ^ (r & 7)) : ci;
        const uint4 pk = *reinterpret_cast<const uint4*>(stage_buf + r * ROW_BYTES + phys * 16);
        const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(w[j]);
          s[2 * j] += f.x; q[2 * j] += f.x * f.x;
          s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        red[(rg * CN + ci * 8 + j) * 2 + 0] = s[j];
        red[(rg * CN + ci * 8 + j) * 2 + 1] = q[j];
      }
      named_bar_sync(bar_id + 1, 128);
      if (e < CN && n0 + c0 + e < p.cout) {
        float ss = 0.f, qq = 0.f;
#pragma unroll 8
        for (int g = 0; g < RG; ++g) {
          ss += red[(g * CN + e) * 2 + 0];
          qq += red[(g * CN + e) * 2 + 1];
        }
        float* dst = p.stats + (static_cast<long long>(img) * p.cout + n0 + c0 + e) * 2;
        atomicAdd(dst, ss);
        atomicAdd(dst + 1, qq);
      }
    }

    // ---- coalesced store of the staged chunk
    const int n_valid = (tile_valid && dbg != 4) ? min(CN, p.cout - n0 - c0) : 0;   // dbg 4: no global stores
    if (n_valid == CN) {
#pragma unroll
      for (int itr = 0; itr < kStoreIters; ++itr) {
        if (row_off[itr] >= 0) {
          const int r = ew * 32 + itr * RPI + lane / CH;
          const int cc = lane % CH;
          const int phys = SWZ ? (cc ^ (r & 7)) : cc;
          const uint4 pk = *reinterpret_cast<const uint4*>(stage_buf + r * ROW_BYTES + phys * 16);
          *reinterpret_cast<uint4*>(out_base + row_off[itr] + n0 + c0 + cc * 8) = pk;
        }
      }
    } else if (n_valid > 0) {
      const int r = e;
      const int wl = r % p.bw, hl = r / p.bw;
      if (w0 + wl < p.w_limit && h0 + hl < p.h_limit) {
        __nv_bfloat16* dst = out_base + img * p.out_sn + (h0 + hl) * p.out_sh +
                             (w0 + wl) * p.out_sw + n0 + c0;
        for (int c = 0; c < n_valid; ++c) {
          const int cc = c >> 3;
          const int phys = SWZ ? (cc ^ (r & 7)) : cc;
          dst[c] = *reinterpret_cast<const __nv_bfloat16*>(stage_buf + r * ROW_BYTES + phys * 16 +
                                                            (c & 7) * 2);
        }
      }
    }
    named_bar_sync(bar_id, 128);   // staging / red buffers are reused by the next chunk
  }
}

}  // namespace b2
