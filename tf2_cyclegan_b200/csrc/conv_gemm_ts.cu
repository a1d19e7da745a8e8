// Tap-shared variant of the cta_group::2 implicit-GEMM convolution (BLOCK_N = 256).
//
// conv_gemm2.cu streams one 128-pixel activation box PER FILTER TAP through shared memory, so a
// 3x3 conv pulls every activation through L2 -> SM nine times; measured with the MMAs removed
// the kernel is no faster, i.e. it is bound by operand delivery (~8 TB/s of the ~12 TB/s L2
// throughput cap), not by the tensor pipe. Here the producer loads ONE halo tile per 64-channel
// block — all pixels any tap of the tile touches — and the MMA thread reads it once per tap
// through a row-shifted K-major descriptor. Two rings: A (halo tiles, 2 stages) and B (weight
// half-tiles per (channel block, tap), 7 stages). K order is (channel block, tap).
//
// Everything else (CTA pair, both CTAs' loads completing on the leader's barriers, multicast
// commits, double-buffered TMEM accumulators, shared epilogue) is as in conv_gemm2.cu.
#include <cuda_runtime.h>

#include "conv_epilogue.cuh"

namespace b2 {

int dbg_flag();

struct TsCfg {
  static constexpr int BLOCK_N = 256;
  static constexpr int kAStageBytes = 34 * 1024;       // 272 rows x 128 B (flat: 2 boxes of 136 rows)
  static constexpr int kAStages = 2;
  static constexpr int kBBytes = 128 * 128;            // this CTA's half of the weight tile (one tap)
  // taps per weight stage. 2 (one wait / commit for eight MMAs, 3 stages) measured the same as 1 (6 stages):
  // with 512 tensor cycles per tap this kernel is not bound by the issue warp, unlike its MXFP8 sibling
  static constexpr int kTaps = 1;
  static constexpr int kBStride = kTaps * kBBytes;
  static constexpr int kBStages = 6;
  static constexpr int kAccCols = BLOCK_N;
  static constexpr int kTmemCols = 2 * kAccCols;
  static constexpr int kHalfN = BLOCK_N / 2;           // columns per epilogue warpgroup
  static constexpr int kStagingBytes = 2 * EpiCfg<kHalfN>::kStagingBytes;
  static constexpr int kRedBytes = 2 * EpiCfg<kHalfN>::kRedBytes;
  static constexpr int kNumBars = 2 * kAStages + 2 * kBStages + 4;
  static constexpr int kBarrierBytes = kNumBars * 8 + 16;
  static constexpr int kSmemBytes = kAStages * kAStageBytes + kBStages * kBStride + kStagingBytes +
                                    kRedBytes + kBarrierBytes + 1024;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <bool B_MN>
__global__ void __launch_bounds__(384, 1)
conv_gemm_ts_kernel(const __grid_constant__ ConvGemmParams p, int num_images, int dbg) {
  using Cfg = TsCfg;
  constexpr int BLOCK_N = Cfg::BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + Cfg::kAStages * Cfg::kAStageBytes;
  uint8_t* stage_buf = b_ring + Cfg::kBStages * Cfg::kBStride;
  float* red = reinterpret_cast<float*>(stage_buf + Cfg::kStagingBytes);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(stage_buf + Cfg::kStagingBytes + Cfg::kRedBytes);
  uint64_t* a_empty = a_full + Cfg::kAStages;
  uint64_t* b_full = a_empty + Cfg::kAStages;
  uint64_t* b_empty = b_full + Cfg::kBStages;
  uint64_t* tmem_full_bar = b_empty + Cfg::kBStages;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();               // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int m_tiles = tiles_per_img * num_images;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int n_tiles = (p.cout + BLOCK_N - 1) / BLOCK_N;
  const int total_pairs = m_pairs * n_tiles;
  const TapShare ts = p.ts;
  // byte offset of every tap's first row inside the halo tile
  __shared__ uint32_t tap_off[kMaxTaps];
  __shared__ int tap_bk[kMaxTaps];
  if (threadIdx.x < p.num_taps) {
    const Tap tap = p.taps[threadIdx.x];
    tap_bk[threadIdx.x] = tap.bk;
    tap_off[threadIdx.x] =
        static_cast<uint32_t>((tap.dh - ts.origin_dh) * ts.pitch + (tap.dw - ts.origin_dw)) * 128u;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmTS);
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kAStages; ++s) {
      mbar_init(&a_full[s], 2);          // one arrival per CTA of the pair (used in the leader)
      mbar_init(&a_empty[s], 1);         // multicast commit from the leader's MMA thread
    }
    for (int s = 0; s < Cfg::kBStages; ++s) {
      mbar_init(&b_full[s], 2);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 512);   // 256 epilogue threads x 2 CTAs (used in the leader)
    }
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 2) tmem_alloc_2sm(tmem_ptr_smem, Cfg::kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  auto decode = [&](int pt, int& mt, int& n0, int& img, int& h0, int& w0) {
    const int mp = pt / n_tiles;
    n0 = (pt - mp * n_tiles) * BLOCK_N;
    mt = 2 * mp + static_cast<int>(rank);
    img = mt / tiles_per_img;            // may be == num_images for the phantom tile: TMA zero-fills
    const int t_in = mt - img * tiles_per_img;
    const int th = t_in / p.tiles_w;
    w0 = (t_in - th * p.tiles_w) * p.bw;
    h0 = th * p.bh;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    // warp-uniform loop, one elected lane issues (keeps tensor-map / barrier operands uniform)
    {
      uint32_t ita = 0, itb = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
        int mt, n0, img, h0, w0;
        decode(pt, mt, n0, img, h0, w0);
        for (int cb = 0; cb < p.cin_blocks; ++cb, ++ita) {
          const int sa = ita % Cfg::kAStages;
          mbar_wait(&a_empty[sa], ((ita / Cfg::kAStages) & 1) ^ 1);
          uint8_t* abuf = a_ring + sa * Cfg::kAStageBytes;
          if (elect_one()) {
            if (dbg == 3) {   // timing experiment: no loads, barriers only
              if (rank == 0) mbar_arrive(&a_full[sa]);
            } else {
            if (rank == 0) mbar_arrive_expect_tx(&a_full[sa], 2 * ts.nbox * ts.box_bytes);
            for (int b = 0; b < ts.nbox; ++b)
              tma_load_4d_2sm(abuf + b * ts.box_bytes, &p.tmTS, &a_full[sa], cb * 64,
                              w0 + ts.origin_dw + b * ts.box_step, h0 + ts.origin_dh, img);
            }
            if (rank != 0) mbar_arrive_cluster(&a_full[sa], 0);
          }
          __syncwarp();
          for (int t = 0; t < p.num_taps; t += Cfg::kTaps, ++itb) {
            const int nk = min(Cfg::kTaps, p.num_taps - t);
            const int sb = itb % Cfg::kBStages;
            mbar_wait(&b_empty[sb], ((itb / Cfg::kBStages) & 1) ^ 1);
            uint8_t* bbuf = b_ring + sb * Cfg::kBStride;
            if (elect_one()) {
              if (dbg == 3) {
                if (rank == 0) mbar_arrive(&b_full[sb]);
              } else {
              if (rank == 0) mbar_arrive_expect_tx(&b_full[sb], 2 * nk * Cfg::kBBytes);
              for (int j = 0; j < nk; ++j) {
                const int bk = tap_bk[t + j];
                uint8_t* bj = bbuf + j * Cfg::kBBytes;
                if constexpr (!B_MN) {
                  tma_load_2d_2sm(bj, &p.tmB, &b_full[sb], bk + cb * 64, n0 + rank * 128);
                } else {
#pragma unroll
                  for (int jj = 0; jj < 2; ++jj)
                    tma_load_2d_2sm(bj + jj * (64 * 128), &p.tmB, &b_full[sb],
                                    bk + n0 + rank * 128 + jj * 64, cb * 64);
                }
              }
              }
              if (rank != 0) mbar_arrive_cluster(&b_full[sb], 0);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // warp-uniform loop (descriptors stay in uniform registers); one elected lane issues
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N, 0, B_MN ? 1 : 0);
      const uint32_t a_ring_u = smem_u32(a_ring), b_ring_u = smem_u32(b_ring);
      uint32_t ita = 0, itb = 0, tile_i = 0;
      // The issue loop is latency-bound on its own instruction stream (measured on the MXFP8 sibling of this
      // kernel, profiles/r2/mx_timeline.txt: ~140 cycles per barrier wait even when the barrier is complete,
      // ~60 per MMA when every tap starts with a load -> address -> R2UR chain). So: the next stage's / block's
      // barriers are probed (non-blocking) BEFORE the current MMAs are issued, the next tap's halo offset is
      // loaded one iteration ahead, and descriptors are built as (constant high word, 14-bit address field).
      bool b_ready = false, a_ready = false;
      const uint32_t da_hi = ((static_cast<uint32_t>(ts.sbo) >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      constexpr uint32_t db_hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      constexpr uint32_t db_lbo = B_MN ? ((64u * 128u) >> 4) << 16 : 0u;     // MN-major: stride between 64-wide atoms
      uint32_t off_next = tap_off[0] >> 4;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters, ++tile_i) {
        const uint32_t acc = tile_i & 1;
        mbar_wait(&tmem_empty_bar[acc], ((tile_i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * Cfg::kAccCols;
        uint32_t accumulate = 0;
        for (int cb = 0; cb < p.cin_blocks; ++cb, ++ita) {
          const int sa = ita % Cfg::kAStages;
          if (!a_ready) mbar_wait(&a_full[sa], (ita / Cfg::kAStages) & 1);
          tc_fence_after();
          const uint32_t a_enc = ((a_ring_u + sa * Cfg::kAStageBytes) >> 4) & 0x3FFFu;
          for (int t = 0; t < p.num_taps; t += Cfg::kTaps, ++itb) {
            const int nk = min(Cfg::kTaps, p.num_taps - t);
            const int sb = itb % Cfg::kBStages;
            const uint32_t off0 = off_next;
            const uint32_t off1 = tap_off[t + 1 < p.num_taps ? t + 1 : 0] >> 4;
            off_next = tap_off[t + Cfg::kTaps < p.num_taps ? t + Cfg::kTaps : 0] >> 4;
            if (!b_ready) mbar_wait(&b_full[sb], (itb / Cfg::kBStages) & 1);
            tc_fence_after();
            {
              const uint32_t nb = itb + 1;
              b_ready = mbar_test_wait(&b_full[nb % Cfg::kBStages], (nb / Cfg::kBStages) & 1);
              if (t + Cfg::kTaps >= p.num_taps) {
                const uint32_t na = ita + 1;
                a_ready = mbar_test_wait(&a_full[na % Cfg::kAStages], (na / Cfg::kAStages) & 1);
              }
            }
            const uint32_t b_enc = ((b_ring_u + sb * Cfg::kBStride) >> 4) & 0x3FFFu;
            if (elect_one()) {
              if (dbg != 2) {
#pragma unroll
                for (int j = 0; j < Cfg::kTaps; ++j) {
                  if (j < nk) {
                    const uint64_t da = (static_cast<uint64_t>(da_hi) << 32) | (a_enc + (j == 0 ? off0 : off1));
                    const uint64_t db = (static_cast<uint64_t>(db_hi) << 32) | ((b_enc + j * (Cfg::kBBytes >> 4)) | db_lbo);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      umma_bf16_2sm(tmem_d, da + static_cast<uint64_t>(k * 2),
                                    db + static_cast<uint64_t>(B_MN ? k * 128 : k * 2), idesc,
                                    (j == 0 && k == 0) ? accumulate : 1u);
                  }
                }
              }
              umma_commit_2sm(&b_empty[sb]);
            }
            __syncwarp();
            accumulate = 1;
          }
          if (elect_one()) umma_commit_2sm(&a_empty[sa]);
          __syncwarp();
        }
        if (elect_one()) umma_commit_2sm(&tmem_full_bar[acc]);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own TMEM half) =====================
    // two warpgroups (warps 4-7, 8-11): each drains half of the accumulator columns
    const int wg = (warp - 4) >> 2;
    const int e = (threadIdx.x - 128) & 127;
    const int ew = warp & 3;
    uint8_t* my_stage = stage_buf + wg * EpiCfg<Cfg::kHalfN>::kStagingBytes;
    float* my_red = red + wg * (EpiCfg<Cfg::kHalfN>::kRedBytes / 4);
    uint32_t tile_i = 0;
    for (int pt = cluster_id; pt < total_pairs; pt += num_clusters, ++tile_i) {
      int mt, n0, img, h0, w0;
      decode(pt, mt, n0, img, h0, w0);
      const uint32_t acc = tile_i & 1;
      mbar_wait(&tmem_full_bar[acc], (tile_i >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * Cfg::kAccCols + wg * Cfg::kHalfN +
                             (static_cast<uint32_t>(ew * 32) << 16);
      if (dbg == 1) { tc_fence_before(); mbar_arrive_cluster(&tmem_empty_bar[acc], 0); continue; }
      conv_epilogue_tile<Cfg::kHalfN>(p, p.out, my_stage, my_red, taddr, img, h0, w0, n0 + wg * Cfg::kHalfN,
                                      e, ew, lane, mt < m_tiles,
                                      [&]() { mbar_arrive_cluster(&tmem_empty_bar[acc], 0); }, dbg, 1 + 2 * wg);
    }
  }

  tc_fence_before();
  cluster_sync_all();                    // the peer may still read this CTA's smem / signal its barriers
  if (warp == 2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
}

template <bool B_MN>
static void launch_ts_t(const ConvGemmParams& p, int num_images, cudaStream_t stream) {
  using Cfg = TsCfg;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(conv_gemm_ts_kernel<B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         Cfg::kSmemBytes);
    configured = true;
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int m_tiles = p.tiles_w * p.tiles_h * num_images;
  const int pairs = ((m_tiles + 1) / 2) * ((p.cout + 255) / 256);
  const int clusters = pairs < sms / 2 ? pairs : sms / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, conv_gemm_ts_kernel<B_MN>, p, num_images, dbg_flag());
}

// false: the launch does not qualify (caller falls back to the per-tap kernels)
bool launch_conv_gemm_ts(const ConvGemmParams& p, int block_n, bool b_mn_major, int num_images,
                         cudaStream_t stream) {
  const TapShare& ts = p.ts;
  if (!ts.enabled || block_n != 256 || p.fp8 || p.num_phases != 1) return false;
  if (ts.nbox < 1 || ts.nbox * ts.box_bytes > TsCfg::kAStageBytes) return false;
  if (ts.nbox > 1 && ts.box_bytes % 1024 != 0) return false;
  if (b_mn_major) launch_ts_t<true>(p, num_images, stream);
  else launch_ts_t<false>(p, num_images, stream);
  return true;
}

}  // namespace b2
