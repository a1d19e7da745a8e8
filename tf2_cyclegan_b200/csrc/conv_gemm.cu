// Implicit-GEMM convolution for sm_100a: TMA (4-D tiled boxes, zero OOB fill) →
// 128B-swizzled shared memory → tcgen05.mma (M=128, N=BLOCK_N, K=16, fp32
// accumulators in TMEM) → tcgen05.ld epilogue with fused bias / activation /
// InstanceNorm partial statistics.
//
//   D[m, n] = sum_taps sum_c A_tap[m, c] * W[n, tap, c]
//
// One CTA computes one 128-pixel x BLOCK_N-channel output tile; the filter taps
// are walked by shifting the TMA box coordinates, so no im2col buffer exists.
// The same kernel serves forward convs (box mode over padded / parity views),
// transposed-conv phases and dgrad (flat mode over zero-ringed gradients); the
// weight operand may be K-major ([N][K]) or MN-major ([K][N], used by dgrad so
// forward and backward share one weight copy).
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer,
// warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM lane quarter = warp % 4).
#include "conv_gemm.cuh"

namespace b2 {

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kABytes = 128 * 128;        // 128 pixels x 64 bf16
  static constexpr int kBBytes = BLOCK_N * 128;    // BLOCK_N x 64 bf16
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N >= 256) ? 4 : (BLOCK_N >= 128 ? 6 : 8);
  static constexpr int kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
  static constexpr int kChunks = BLOCK_N / 8;      // 16-byte chunks per staged output row
  static constexpr int kRowGroups = 128 / (kChunks > 128 ? 128 : kChunks);
  static constexpr int kStagingBytes = 128 * BLOCK_N * 2;
  static constexpr int kRedBytes = kRowGroups * BLOCK_N * 2 * 4;
  static constexpr int kBarrierBytes = (2 * kStages + 1) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + 1024;  // +align slack
  static_assert(kStagingBytes + kRedBytes <= kStages * kStageBytes, "epilogue aliases the pipeline");
};

template <int BLOCK_N, bool B_MN>
__global__ void __launch_bounds__(256, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile -> (image, h0, w0)
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int img = blockIdx.x / tiles_per_img;
  const int t_in = blockIdx.x - img * tiles_per_img;
  const int th = t_in / p.tiles_w;
  const int w0 = (t_in - th * p.tiles_w) * p.bw;
  const int h0 = th * p.bh;
  const int n0 = blockIdx.y * BLOCK_N;
  const int num_kb = p.num_taps * p.cin_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmA[0]);
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int kb = 0;
      for (int t = 0; t < p.num_taps; ++t) {
        const Tap tap = p.taps[t];
        const CUtensorMap* amap = &p.tmA[tap.map];
        for (int cb = 0; cb < p.cin_blocks; ++cb, ++kb) {
          const int s = kb % Cfg::kStages;
          const uint32_t ph = (kb / Cfg::kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
          tma_load_4d(sa, amap, &full_bar[s], cb * 64, w0 + tap.dw, h0 + tap.dh, img + tap.dn);
          if constexpr (!B_MN) {
            tma_load_2d(sb, &p.tmB, &full_bar[s], tap.bk + cb * 64, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(sb + j * (64 * 128), &p.tmB, &full_bar[s], tap.bk + n0 + j * 64, cb * 64);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================= MMA issuer =============================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, B_MN ? 1 : 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % Cfg::kStages;
        const uint32_t ph = (kb / Cfg::kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t sb = sa + Cfg::kABytes;
        const uint64_t da = make_desc_kmajor(sa);
        const uint64_t db = B_MN ? make_desc_mnmajor(sb, 64) : make_desc_kmajor(sb);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // K advance of 16 elements: +32 B inside the swizzle row (K-major),
          // +16 rows * 128 B (MN-major); descriptor address field is in 16-byte units
          const uint64_t da_k = da + static_cast<uint64_t>(k * 2);
          const uint64_t db_k = db + static_cast<uint64_t>(B_MN ? k * 128 : k * 2);
          umma_bf16(tmem_base, da_k, db_k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);   // frees the smem slot when these MMAs retire
      }
      umma_commit(tmem_full_bar);     // accumulator complete
    }
  } else if (warp >= 4) {
    // ============================== epilogue ==============================
    const int e = threadIdx.x - 128;          // 0..127 == tile row == TMEM lane
    const int ew = warp - 4;                  // TMEM lane quarter
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    uint8_t* stage = smem;                    // aliases the (now idle) pipeline buffers
    float* red = reinterpret_cast<float*>(smem + Cfg::kStagingBytes);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    constexpr int ROW_BYTES = BLOCK_N * 2;

    if constexpr (BLOCK_N >= 32) {
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(r[j]);
          if (p.bias != nullptr && n0 + c + j < p.cout) x += __ldg(p.bias + n0 + c + j);
          v[j] = apply_act(x, p.act);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ci = (c >> 3) + q;
          const int phys = ci ^ (e & 7);
          uint4 pk;
          pk.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
          pk.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
          pk.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
          pk.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
          *reinterpret_cast<uint4*>(stage + e * ROW_BYTES + phys * 16) = pk;
        }
      }
    } else {
      uint32_t r[16];
      tmem_ld_32x16(taddr, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x = __uint_as_float(r[j]);
        if (p.bias != nullptr && n0 + j < p.cout) x += __ldg(p.bias + n0 + j);
        v[j] = apply_act(x, p.act);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint4 pk;
        pk.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
        pk.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
        pk.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
        pk.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
        *reinterpret_cast<uint4*>(stage + e * ROW_BYTES + q * 16) = pk;
      }
    }
    tc_fence_before();
    named_bar_sync(1, 128);

    constexpr int CH = Cfg::kChunks;
    constexpr bool SWZ = (CH >= 8);
    // ---- InstanceNorm partial statistics (sum, sum of squares per channel)
    if (p.stats != nullptr) {
      constexpr int RG = Cfg::kRowGroups;
      const int ci = e % CH;
      const int rg = e / CH;
      float s[8], q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
      for (int r = rg; r < 128; r += RG) {
        const int phys = SWZ ? (ci ^ (r & 7)) : ci;
        const uint4 pk = *reinterpret_cast<const uint4*>(stage + r * ROW_BYTES + phys * 16);
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
        red[(rg * BLOCK_N + ci * 8 + j) * 2 + 0] = s[j];
        red[(rg * BLOCK_N + ci * 8 + j) * 2 + 1] = q[j];
      }
      named_bar_sync(1, 128);
      for (int col = e; col < BLOCK_N; col += 128) {
        if (n0 + col < p.cout) {
          float ss = 0.f, qq = 0.f;
#pragma unroll 4
          for (int g = 0; g < RG; ++g) {
            ss += red[(g * BLOCK_N + col) * 2 + 0];
            qq += red[(g * BLOCK_N + col) * 2 + 1];
          }
          float* dst = p.stats + (static_cast<long long>(img) * p.cout + n0 + col) * 2;
          atomicAdd(dst, ss);
          atomicAdd(dst + 1, qq);
        }
      }
    }
    // ---- coalesced store of the staged tile
    const int n_valid = min(BLOCK_N, p.cout - n0);
    if (n_valid == BLOCK_N) {
      constexpr int CPR = CH > 32 ? 32 : CH;     // chunks handled per row per pass
      constexpr int RPI = 32 / CPR;              // rows per warp instruction
      for (int it = 0; it < 32 / RPI; ++it) {
        const int r = ew * 32 + it * RPI + lane / CPR;
        const int wl = r % p.bw, hl = r / p.bw;
        if (w0 + wl < p.w_limit) {
          __nv_bfloat16* dst = p.out + img * p.out_sn + (h0 + hl) * p.out_sh +
                               (w0 + wl) * p.out_sw + n0;
#pragma unroll
          for (int cc = lane % CPR; cc < CH; cc += CPR) {
            const int phys = SWZ ? (cc ^ (r & 7)) : cc;
            const uint4 pk = *reinterpret_cast<const uint4*>(stage + r * ROW_BYTES + phys * 16);
            *reinterpret_cast<uint4*>(dst + cc * 8) = pk;
          }
        }
      }
    } else {
      const int r = e;
      const int wl = r % p.bw, hl = r / p.bw;
      if (w0 + wl < p.w_limit) {
        __nv_bfloat16* dst = p.out + img * p.out_sn + (h0 + hl) * p.out_sh +
                             (w0 + wl) * p.out_sw + n0;
        for (int c = 0; c < n_valid; ++c) {
          const int cc = c >> 3;
          const int phys = SWZ ? (cc ^ (r & 7)) : cc;
          dst[c] = *reinterpret_cast<const __nv_bfloat16*>(stage + r * ROW_BYTES + phys * 16 +
                                                            (c & 7) * 2);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int BLOCK_N, bool B_MN>
static void launch_t(const ConvGemmParams& p, int num_images, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, B_MN>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    configured = true;
  }
  dim3 grid(p.tiles_w * p.tiles_h * num_images, (p.cout + BLOCK_N - 1) / BLOCK_N);
  conv_gemm_kernel<BLOCK_N, B_MN><<<grid, 256, Cfg::kSmemBytes, stream>>>(p);
}

void launch_conv_gemm(const ConvGemmParams& p, int block_n, bool b_mn_major, int num_images,
                      cudaStream_t stream) {
  if (!b_mn_major) {
    switch (block_n) {
      case 16: return launch_t<16, false>(p, num_images, stream);
      case 64: return launch_t<64, false>(p, num_images, stream);
      case 128: return launch_t<128, false>(p, num_images, stream);
      case 256: return launch_t<256, false>(p, num_images, stream);
    }
  } else {
    switch (block_n) {
      case 64: return launch_t<64, true>(p, num_images, stream);
      case 128: return launch_t<128, true>(p, num_images, stream);
      case 256: return launch_t<256, true>(p, num_images, stream);
    }
  }
  fprintf(stderr, "launch_conv_gemm: unsupported block_n=%d mn=%d\n", block_n, int(b_mn_major));
  abort();
}

}  // namespace b2
