// 2-CTA (cta_group::2) variant of the persistent implicit-GEMM convolution: a CTA pair on one
// TPC computes a 256-pixel x BLOCK_N tile with ONE tcgen05.mma stream (M = 256). Each CTA loads
// its own 128-pixel A box and only HALF of the weight tile; the tensor core reads the other
// half from the peer's shared memory, which cuts the L2 -> SM operand traffic per FLOP by 1.5x
// (the 1-CTA kernel is L2-bandwidth bound on the 256-channel layers) while keeping the
// double-buffered TMEM accumulators and the overlapped epilogue.
//
// Protocol (as in CUTLASS / DeepGEMM 2-SM kernels): both CTAs' TMA loads complete on the
// LEADER's full barrier (peer-bit-cleared mbarrier address, cta_group::2 TMA); the leader's MMA
// thread commits with multicast to BOTH CTAs' empty / tmem-full barriers; epilogue threads of
// both CTAs arrive on the leader's tmem-empty barrier; TMEM is allocated with cta_group::2.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only), warp 2 =
// TMEM allocator, warps 4-7 and 8-11 = two epilogue warpgroups, each draining half of the
// accumulator columns (TMEM lane quarter = warp % 4). The producer / MMA loops are executed by
// the whole warp with only the issue instructions under elect.sync: inside an `if (lane == 0)`
// region the compiler wraps every uniform operand of UTMALDG / UTCHMMA in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop and the single issuing thread becomes the bottleneck.
#include <cuda_runtime.h>

#include "conv_epilogue.cuh"

namespace b2 {

int dbg_flag();

template <int BLOCK_N>
struct Gemm2Cfg {
  static constexpr int kHalfN = BLOCK_N / 2;
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = kHalfN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N >= 256) ? 5 : 7;
  static constexpr int kAccCols = BLOCK_N;
  static constexpr int kTmemCols = 2 * kAccCols;
  // two epilogue warpgroups, each owns half of the columns and its own staging / red buffers
  static constexpr int kStagingBytes = 2 * EpiCfg<kHalfN>::kStagingBytes;
  static constexpr int kRedBytes = 2 * EpiCfg<kHalfN>::kRedBytes;
  static constexpr int kBarrierBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kStagingBytes + kRedBytes + kBarrierBytes + 1024;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int BLOCK_N, bool B_MN, bool FP8>
__global__ void __launch_bounds__(384, 1)
conv_gemm2_kernel(const __grid_constant__ ConvGemmParams p, int num_images, int dbg) {
  using Cfg = Gemm2Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* stage_buf = smem + Cfg::kStages * Cfg::kStageBytes;
  float* red = reinterpret_cast<float*>(stage_buf + Cfg::kStagingBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_buf + Cfg::kStagingBytes + Cfg::kRedBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;    // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  constexpr int KB = FP8 ? 128 : 64;   // elements per 128-byte K block
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();               // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int m_tiles = tiles_per_img * num_images;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int n_tiles = (p.cout + BLOCK_N - 1) / BLOCK_N;
  const int pairs_per_phase = m_pairs * n_tiles;
  const int total_pairs = pairs_per_phase * p.num_phases;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmA[0]);
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 2);        // one arrival per CTA of the pair (used in the leader)
      mbar_init(&empty_bar[s], 1);       // multicast commit from the leader's MMA thread
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 512);   // 256 epilogue threads x 2 CTAs (used in the leader)
    }
    fence_barrier_init();
  }
  cluster_sync_all();                    // barriers of both CTAs are initialised
  if (warp == 2) tmem_alloc_2sm(tmem_ptr_smem, Cfg::kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // pair-tile -> this CTA's 128-pixel tile
  auto decode = [&](int ptile, int& mt, int& n0, int& img, int& h0, int& w0) {
    const int pt = ptile % pairs_per_phase;
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
      uint32_t it = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
        int mt, n0, img, h0, w0;
        decode(pt, mt, n0, img, h0, w0);
        const int phase = pt / pairs_per_phase;
        const int t_begin = p.phase_tap_begin[phase], t_end = t_begin + p.phase_tap_count[phase];
        for (int t = t_begin; t < t_end; ++t) {
          const Tap tap = p.taps[t];
          const CUtensorMap* amap = &p.tmA[tap.map];
          for (int cb = 0; cb < p.cin_blocks; ++cb, ++it) {
            const int s = it % Cfg::kStages;
            const uint32_t ph = (it / Cfg::kStages) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* sa = smem + s * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            if (elect_one()) {
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::kStageBytes);
              tma_load_4d_2sm(sa, amap, &full_bar[s], cb * KB, w0 + tap.dw, h0 + tap.dh, img + tap.dn);
              if constexpr (!B_MN) {
                tma_load_2d_2sm(sb, &p.tmB, &full_bar[s], tap.bk + cb * KB, n0 + rank * Cfg::kHalfN);
              } else {
#pragma unroll
                for (int j = 0; j < Cfg::kHalfN / 64; ++j)
                  tma_load_2d_2sm(sb + j * (64 * 128), &p.tmB, &full_bar[s],
                                  tap.bk + n0 + rank * Cfg::kHalfN + j * 64, cb * KB);
              }
              if (rank != 0) mbar_arrive_cluster(&full_bar[s], 0);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // The whole warp runs the loop (warp-uniform control flow and operands, so descriptors
    // live in uniform registers); one elected lane issues the MMAs and commits.
    if (rank == 0) {
      constexpr uint32_t idesc = FP8 ? make_idesc_e4m3(256, BLOCK_N) : make_idesc_bf16(256, BLOCK_N, 0, B_MN ? 1 : 0);
      const uint32_t smem_base = smem_u32(smem);
      uint32_t it = 0, tile_i = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters, ++tile_i) {
        const int num_kb = p.phase_tap_count[pt / pairs_per_phase] * p.cin_blocks;
        const uint32_t acc = tile_i & 1;
        const uint32_t acc_ph = (tile_i >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * Cfg::kAccCols;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t da = make_desc_kmajor(sa);
          const uint64_t db = B_MN ? make_desc_mnmajor(sb, 64) : make_desc_kmajor(sb);
          if (elect_one()) {
            if (dbg != 2) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da_k = da + static_cast<uint64_t>(k * 2);
                const uint64_t db_k = db + static_cast<uint64_t>(B_MN ? k * 128 : k * 2);
                if constexpr (FP8) umma_f8_2sm(tmem_d, da_k, db_k, idesc, (kb | k) != 0 ? 1u : 0u);
                else umma_bf16_2sm(tmem_d, da_k, db_k, idesc, (kb | k) != 0 ? 1u : 0u);
              }
            }
            umma_commit_2sm(&empty_bar[s]);
          }
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
      const uint32_t acc_ph = (tile_i >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * Cfg::kAccCols + wg * Cfg::kHalfN +
                             (static_cast<uint32_t>(ew * 32) << 16);
      if (dbg == 1) { tc_fence_before(); mbar_arrive_cluster(&tmem_empty_bar[acc], 0); continue; }
      conv_epilogue_tile<Cfg::kHalfN>(p, p.out + p.phase_out_off[pt / pairs_per_phase], my_stage, my_red, taddr,
                                      img, h0, w0, n0 + wg * Cfg::kHalfN, e, ew, lane, mt < m_tiles,
                                      [&]() { mbar_arrive_cluster(&tmem_empty_bar[acc], 0); }, dbg, 1 + 2 * wg);
    }
  }

  tc_fence_before();
  cluster_sync_all();                    // the peer may still read this CTA's smem / signal its barriers
  if (warp == 2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
}

static int num_sms2() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BLOCK_N, bool B_MN, bool FP8 = false>
static void launch2_t(const ConvGemmParams& p, int num_images, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BLOCK_N>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(conv_gemm2_kernel<BLOCK_N, B_MN, FP8>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    configured = true;
  }
  const int m_tiles = p.tiles_w * p.tiles_h * num_images;
  const int pairs = ((m_tiles + 1) / 2) * ((p.cout + BLOCK_N - 1) / BLOCK_N) * p.num_phases;
  const int max_clusters = num_sms2() / 2;
  const int clusters = pairs < max_clusters ? pairs : max_clusters;
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
  cudaLaunchKernelEx(&cfg, conv_gemm2_kernel<BLOCK_N, B_MN, FP8>, p, num_images, dbg_flag());
}

bool launch_conv_gemm_2cta(const ConvGemmParams& p, int block_n, bool b_mn_major, int num_images,
                           cudaStream_t stream) {
  if (p.fp8) {
    if (block_n != 256 || b_mn_major) return false;
    launch2_t<256, false, true>(p, num_images, stream);
    return true;
  }
  if (block_n == 256) {
    if (b_mn_major) launch2_t<256, true>(p, num_images, stream);
    else launch2_t<256, false>(p, num_images, stream);
    return true;
  }
  if (block_n == 128) {
    if (b_mn_major) launch2_t<128, true>(p, num_images, stream);
    else launch2_t<128, false>(p, num_images, stream);
    return true;
  }
  return false;
}

}  // namespace b2
