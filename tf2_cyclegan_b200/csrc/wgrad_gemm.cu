// Weight-gradient GEMM for sm_100a (tcgen05, both operands MN-major, split-K).
//
//   dW[co, tap, ci] += sum_pixels dY[pixel, co] * A_tap[pixel, ci]
//
// The reduction dimension is the pixel axis, which is the *strided* axis of the
// NHWC tensors, so both MMA operands are MN-major: the TMA boxes {64 channels,
// 64 pixels} land in shared memory exactly as the canonical SWIZZLE_128B
// MN-major atoms. One CTA owns (co tile of 128) x (ci tile <= 256) x (one tap)
// x (one K split); partial sums are reduced into the fp32 gradient buffer with
// vector red.global.add, which also accumulates the three calls of a generator.
#include <stdlib.h>

#include "wgrad_gemm.cuh"

namespace b2 {

int dbg_flag();

template <int BLOCK_N>
struct WgradCfg {
  static constexpr int kABytes = 2 * 64 * 128;          // 128 co x 64 pixels (2 MN atoms)
  static constexpr int kBBytes = (BLOCK_N / 64) * 64 * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N >= 256) ? 4 : (BLOCK_N >= 128 ? 6 : 8);
  static constexpr int kTmemCols = BLOCK_N;
  static constexpr int kBarrierBytes = (2 * kStages + 1) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + 1024;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b),
               "f"(c), "f"(d)
               : "memory");
}


// Epilogue shared by both wgrad kernels: TMEM row (one output channel) -> padded smem row ->
// ONE bulk fp32 reduce-add per row into the gradient buffer (TMA engine, no per-thread atomics).
// `stage` aliases the (idle) pipeline buffers; rows are padded by 16 B against bank conflicts.
template <int BLOCK_N>
__device__ __forceinline__ void wgrad_epilogue(const WgradParams& p, uint8_t* stage, uint32_t taddr,
                                               int e, int co, int n0, int col0) {
  constexpr int ROW_BYTES = BLOCK_N * 4 + 16;
  float* srow = reinterpret_cast<float*>(stage + e * ROW_BYTES);
#pragma unroll 1
  for (int c = 0; c < BLOCK_N; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<uint4*>(srow + c + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
  }
  fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the bulk engine
  const int n_valid = min(BLOCK_N, p.cin - n0);
  if (co < p.cout && n_valid > 0) {
    float* grow = p.dw + static_cast<long long>(co) * p.ldw + col0 + n0;
    bulk_reduce_add_f32(grow, srow, static_cast<uint32_t>(n_valid) * 4u);
    tma_store_commit();
    tma_store_wait_all();              // smem row must stay intact until the engine has read it
  }
}

template <int BLOCK_N>
__global__ void __launch_bounds__(256, 1)
wgrad_gemm_kernel(const __grid_constant__ WgradParams p, int dbg) {
  using Cfg = WgradCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.cin + BLOCK_N - 1) / BLOCK_N;
  const int m0 = (blockIdx.x / n_tiles) * 128;
  const int n0 = (blockIdx.x % n_tiles) * BLOCK_N;
  const Tap tap = p.taps[blockIdx.y];
  const int ktiles_per_img = p.ktiles_w * p.ktiles_h;
  const int total_kt = ktiles_per_img * p.num_images;
  // contiguous K range of this split
  const int per = (total_kt + gridDim.z - 1) / gridDim.z;
  const int kt_begin = blockIdx.z * per;
  const int kt_end = min(total_kt, kt_begin + per);
  const int num_kb = kt_end - kt_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmDY[tap.dn]);
    tma_prefetch_desc(&p.tmA[tap.map]);
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

  if (num_kb > 0) {
    // producer and MMA warps run warp-uniform loops (operands stay in uniform registers, the K
    // tile coordinates advance incrementally); one elected lane issues
    if (warp == 0) {
      const CUtensorMap* amap = &p.tmA[tap.map];
      int img = kt_begin / ktiles_per_img;
      int t_in = kt_begin - img * ktiles_per_img;
      int th = t_in / p.ktiles_w;
      int tw = t_in - th * p.ktiles_w;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int w0 = tw * p.kbw, h0 = th * p.kbh;
        const int s = kb % Cfg::kStages;
        const uint32_t ph = (kb / Cfg::kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
          // channels beyond the tensor extent are zero-filled by TMA (cout / cin < tile)
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_4d(sa + j * (64 * 128), &p.tmDY[tap.dn], &full_bar[s], m0 + j * 64, w0, h0, img);
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_4d(sb + j * (64 * 128), amap, &full_bar[s], n0 + j * 64, w0 + tap.dw,
                        h0 + tap.dh, img);
        }
        __syncwarp();
        if (++tw == p.ktiles_w) { tw = 0; if (++th == p.ktiles_h) { th = 0; ++img; } }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 1, 1);
      const uint32_t smem_base = smem_u32(smem);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % Cfg::kStages;
        const uint32_t ph = (kb / Cfg::kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kABytes;
        const uint64_t da = make_desc_mnmajor(sa, 64);
        const uint64_t db = make_desc_mnmajor(sb, 64);
        if (elect_one()) {
          if (dbg != 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k)   // 16 pixels per MMA = 16 rows x 128 B = 128 x 16-byte units
              umma_bf16(tmem_base, da + static_cast<uint64_t>(k * 128),
                        db + static_cast<uint64_t>(k * 128), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(tmem_full_bar);
      __syncwarp();
    } else if (warp >= 4) {
      const int e = threadIdx.x - 128;
      const int ew = warp - 4;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
      if (dbg != 1) wgrad_epilogue<BLOCK_N>(p, smem, taddr, e, m0 + e, n0, tap.bk);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int BLOCK_N>
static void launch_t(const WgradParams& p, int ksplit, cudaStream_t stream) {
  using Cfg = WgradCfg<BLOCK_N>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(wgrad_gemm_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         Cfg::kSmemBytes);
    configured = true;
  }
  const int m_tiles = (p.cout + 127) / 128;
  const int n_tiles = (p.cin + BLOCK_N - 1) / BLOCK_N;
  dim3 grid(m_tiles * n_tiles, p.num_taps, ksplit);
  wgrad_gemm_kernel<BLOCK_N><<<grid, 256, Cfg::kSmemBytes, stream>>>(p, dbg_flag());
}


// ---------------------------------------------------------------------------------------
// 2-CTA (cta_group::2) variant: a CTA pair accumulates dW for 256 output channels x BLOCK_N
// input channels of one tap / K split with M = 256 MMAs. Each CTA loads its own 128-channel
// half of dY and HALF of the activation tile (the tensor core reads the rest from the peer's
// shared memory): 32 KB instead of 48 KB of L2 traffic per 64-pixel K tile per SM.
template <int BLOCK_N>
struct Wgrad2Cfg {
  static constexpr int kHalfN = BLOCK_N / 2;
  static constexpr int kABytes = 2 * 64 * 128;
  static constexpr int kBBytes = (kHalfN / 64) * 64 * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = 6;
  static constexpr int kTmemCols = BLOCK_N;
  static constexpr int kBarrierBytes = (2 * kStages + 1) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + 1024;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(256, 1)
wgrad_gemm2_kernel(const __grid_constant__ WgradParams p, int dbg) {
  using Cfg = Wgrad2Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_tiles = (p.cin + BLOCK_N - 1) / BLOCK_N;
  const int pair = blockIdx.x >> 1;                       // (m pair, n tile)
  const int m0 = (pair / n_tiles) * 256 + static_cast<int>(rank) * 128;
  const int n0 = (pair % n_tiles) * BLOCK_N;
  const Tap tap = p.taps[blockIdx.y];
  const int ktiles_per_img = p.ktiles_w * p.ktiles_h;
  const int total_kt = ktiles_per_img * p.num_images;
  const int per = (total_kt + gridDim.z - 1) / gridDim.z;
  const int kt_begin = blockIdx.z * per;
  const int kt_end = min(total_kt, kt_begin + per);
  const int num_kb = kt_end - kt_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmDY[tap.dn]);
    tma_prefetch_desc(&p.tmA[tap.map]);
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 2);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 2) tmem_alloc_2sm(tmem_ptr_smem, Cfg::kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (num_kb > 0) {
    if (warp == 0) {
      const CUtensorMap* amap = &p.tmA[tap.map];
      int img = kt_begin / ktiles_per_img;
      int t_in = kt_begin - img * ktiles_per_img;
      int th = t_in / p.ktiles_w;
      int tw = t_in - th * p.ktiles_w;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int w0 = tw * p.kbw, h0 = th * p.kbh;
        const int s = kb % Cfg::kStages;
        const uint32_t ph = (kb / Cfg::kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::kStageBytes);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_4d_2sm(sa + j * (64 * 128), &p.tmDY[tap.dn], &full_bar[s], m0 + j * 64, w0, h0, img);
#pragma unroll
          for (int j = 0; j < Cfg::kHalfN / 64; ++j)
            tma_load_4d_2sm(sb + j * (64 * 128), amap, &full_bar[s],
                            n0 + static_cast<int>(rank) * Cfg::kHalfN + j * 64, w0 + tap.dw, h0 + tap.dh,
                            img);
          if (rank != 0) mbar_arrive_cluster(&full_bar[s], 0);
        }
        __syncwarp();
        if (++tw == p.ktiles_w) { tw = 0; if (++th == p.ktiles_h) { th = 0; ++img; } }
      }
    } else if (warp == 1) {
      if (rank == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N, 1, 1);
        const uint32_t smem_base = smem_u32(smem);
        bool ready = false;   // non-blocking probe of the next stage, issued before the current MMAs
        for (int kb = 0; kb < num_kb; ++kb) {
          const int s = kb % Cfg::kStages;
          const uint32_t ph = (kb / Cfg::kStages) & 1;
          if (!ready) mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          ready = mbar_test_wait(&full_bar[(kb + 1) % Cfg::kStages], ((kb + 1) / Cfg::kStages) & 1);
          const uint32_t sa = smem_base + s * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t da = make_desc_mnmajor(sa, 64);
          const uint64_t db = make_desc_mnmajor(sb, 64);
          if (elect_one()) {
            if (dbg != 2) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_2sm(tmem_base, da + static_cast<uint64_t>(k * 128),
                              db + static_cast<uint64_t>(k * 128), idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[s]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit_2sm(tmem_full_bar);
        __syncwarp();
      }
    } else if (warp >= 4) {
      const int e = threadIdx.x - 128;
      const int ew = warp - 4;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
      if (dbg != 1) wgrad_epilogue<BLOCK_N>(p, smem, taddr, e, m0 + e, n0, tap.bk);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
}

template <int BLOCK_N>
static void launch2_t(const WgradParams& p, int ksplit, cudaStream_t stream) {
  using Cfg = Wgrad2Cfg<BLOCK_N>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(wgrad_gemm2_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         Cfg::kSmemBytes);
    configured = true;
  }
  const int m_pairs = (p.cout + 255) / 256;
  const int n_tiles = (p.cin + BLOCK_N - 1) / BLOCK_N;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * m_pairs * n_tiles, p.num_taps, ksplit);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, wgrad_gemm2_kernel<BLOCK_N>, p, dbg_flag());
}

bool use_2cta();   // conv_gemm.cu: process-wide kernel options
static bool wgrad_use_2cta() { return use_2cta(); }

bool launch_wgrad_multi(const WgradParams& p, int ksplit, cudaStream_t stream);

void launch_wgrad_gemm(const WgradParams& p, int block_n, int ksplit, cudaStream_t stream) {
  if (ksplit < 0) {   // negative ksplit: caller asks for the multi-tap kernel (all taps in one CTA)
    if (launch_wgrad_multi(p, -ksplit, stream)) return;
    ksplit = -ksplit;
  }
  if (wgrad_use_2cta() && p.cout >= 256 && p.cout % 256 == 0 && block_n >= 128) {
    // keep the CTA count: a pair does the work of two single CTAs
    if (block_n == 256) return launch2_t<256>(p, ksplit, stream);
    return launch2_t<128>(p, ksplit, stream);
  }
  switch (block_n) {
    case 64: return launch_t<64>(p, ksplit, stream);
    case 128: return launch_t<128>(p, ksplit, stream);
    case 256: return launch_t<256>(p, ksplit, stream);
  }
  fprintf(stderr, "launch_wgrad_gemm: unsupported block_n=%d\n", block_n);
  abort();
}

}  // namespace b2
