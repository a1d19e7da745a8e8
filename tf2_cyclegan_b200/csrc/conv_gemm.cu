// Persistent implicit-GEMM convolution for sm_100a: TMA (4-D tiled boxes, zero OOB
// fill) → 128B-swizzled shared memory ring → tcgen05.mma (M=128, N=BLOCK_N, K=16,
// fp32 accumulators in TMEM, double-buffered) → tcgen05.ld epilogue with fused bias /
// activation / InstanceNorm partial statistics, overlapped with the next tile's MMAs.
//
//   D[m, n] = sum_taps sum_c A_tap[m, c] * W[n, tap, c]
//
// One CTA per SM loops over 128-pixel x BLOCK_N-channel output tiles; filter taps are
// walked by shifting the TMA box coordinates, so no im2col buffer exists. The same
// kernel serves forward convs (box mode over padded / parity views), transposed-conv
// phases and dgrad (flat mode over zero-ringed gradients); the weight operand may be
// K-major ([N][K]) or MN-major ([K][N], used by dgrad so forward and backward share one
// weight copy).
//
// Warp roles (256 threads): warp 0 = TMA producer (runs ahead across tiles), warp 1 =
// MMA issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM lane quarter = warp % 4).
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty x2 (MMA <-> epilogue).
#include <stdlib.h>

#include "conv_epilogue.cuh"

namespace b2 {

// Kernel options of this process (set once from tf2_cyclegan_b200/config.py through
// _C.set_kernel_options; no environment reads here). dbg != 0 are timing experiments.
struct KernelOptions {
  int dbg = 0;          // 1 skip epilogue, 2 skip MMAs, 3 skip TMA loads, 4 skip global stores
  bool two_cta = true;  // cta_group::2 kernels for wide layers
  int stages = 0;       // cap on the 1-CTA kernel's smem ring depth
  int grid = 0;         // cap on the 1-CTA kernel's persistent grid
};
static KernelOptions g_opt;
int dbg_flag();

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kABytes = 128 * 128;        // 128 pixels x 64 bf16
  static constexpr int kBBytes = BLOCK_N * 128;    // BLOCK_N x 64 bf16
  static constexpr int kKbBytes = kABytes + kBBytes;              // one K block (64 channels of one tap)
  // K blocks per pipeline stage. The MMA warp's own instruction stream costs ~800 cycles per barrier
  // round trip (measured: profiles/r2/stem_probe.log) against 32 x N / 64 cycles of tensor work per K
  // block, so narrow tiles batch two K blocks per stage (one wait, one commit for eight MMAs).
  static constexpr int kKB = (BLOCK_N <= 128) ? 2 : 1;
  static constexpr int kStageBytes = kKB * kKbBytes;
  static constexpr int kStages = (BLOCK_N >= 256) ? 4 : (BLOCK_N >= 128 ? 3 : (BLOCK_N >= 64 ? 4 : 5));
  static constexpr int kAccCols = BLOCK_N < 32 ? 32 : BLOCK_N;
  static constexpr int kTmemCols = 2 * kAccCols;   // 64 .. 512, power of two
  static constexpr int kChunkN = BLOCK_N < 64 ? BLOCK_N : 64;   // epilogue column chunk
  static constexpr int kChunks16 = kChunkN / 8;    // 16-byte pieces per staged row
  static constexpr int kRowGroups = 128 / kChunks16;
  static constexpr int kStagingBytes = 128 * kChunkN * 2;
  static constexpr int kRedBytes = kRowGroups * kChunkN * 2 * 4;
  static constexpr int kBarrierBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kStagingBytes + kRedBytes + kBarrierBytes + 1024;  // +align slack
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int BLOCK_N, bool B_MN>
__global__ void __launch_bounds__(256, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p, int num_images, int nstages, int dbg) {
  using Cfg = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* stage_buf = smem + Cfg::kStages * Cfg::kStageBytes;
  float* red = reinterpret_cast<float*>(stage_buf + Cfg::kStagingBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_buf + Cfg::kStagingBytes + Cfg::kRedBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;    // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int m_tiles = tiles_per_img * num_images;
  const int n_tiles = (p.cout + BLOCK_N - 1) / BLOCK_N;
  const int tiles_per_phase = m_tiles * n_tiles;
  const int total_tiles = tiles_per_phase * p.num_phases;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmA[0]);
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 128);
    }
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
    // warp-uniform loop; one elected lane issues (operands stay in uniform registers)
    {
      uint32_t s = 0, ph = 0;   // ring slot / phase of the stage being filled
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int phase = tile / tiles_per_phase;
        const int tp = tile - phase * tiles_per_phase;
        const int mt = tp / n_tiles;
        const int n0 = (tp - mt * n_tiles) * BLOCK_N;
        const int img = mt / tiles_per_img;
        const int t_in = mt - img * tiles_per_img;
        const int th = t_in / p.tiles_w;
        const int w0 = (t_in - th * p.tiles_w) * p.bw;
        const int h0 = th * p.bh;
        const int t_begin = p.phase_tap_begin[phase], t_end = t_begin + p.phase_tap_count[phase];
        const int num_kb = p.phase_tap_count[phase] * p.cin_blocks;
        int kb = 0;
        for (int t = t_begin; t < t_end; ++t) {
          const Tap tap = p.taps[t];
          const CUtensorMap* amap = &p.tmA[tap.map];
          for (int cb = 0; cb < p.cin_blocks; ++cb, ++kb) {
            const int j = kb % Cfg::kKB;                      // position inside the stage
            if (j == 0) mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* sa = smem + s * Cfg::kStageBytes + j * Cfg::kKbBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            if (elect_one()) {
              if (dbg == 3) {
                if (j == 0) mbar_arrive(&full_bar[s]);
              } else {
                if (j == 0) mbar_arrive_expect_tx(&full_bar[s], min(Cfg::kKB, num_kb - kb) * Cfg::kKbBytes);
                tma_load_4d(sa, amap, &full_bar[s], cb * 64, w0 + tap.dw, h0 + tap.dh, img + tap.dn);
                if constexpr (!B_MN) {
                  if constexpr (BLOCK_N >= 128) {   // the host encodes half-height boxes for wide tiles
                    tma_load_2d(sb, &p.tmB, &full_bar[s], tap.bk + cb * 64, n0);
                    tma_load_2d(sb + (BLOCK_N / 2) * 128, &p.tmB, &full_bar[s], tap.bk + cb * 64,
                                n0 + BLOCK_N / 2);
                  } else {
                    tma_load_2d(sb, &p.tmB, &full_bar[s], tap.bk + cb * 64, n0);
                  }
                } else {
#pragma unroll
                  for (int jj = 0; jj < BLOCK_N / 64; ++jj)
                    tma_load_2d(sb + jj * (64 * 128), &p.tmB, &full_bar[s], tap.bk + n0 + jj * 64, cb * 64);
                }
              }
            }
            __syncwarp();
            if (j == Cfg::kKB - 1 || kb == num_kb - 1) {      // stage complete
              if (++s == static_cast<uint32_t>(nstages)) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================= MMA issuer =============================
    // warp-uniform loop; one elected lane issues the MMAs and commits
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, B_MN ? 1 : 0);
      const uint32_t smem_base = smem_u32(smem);
      uint32_t s = 0, ph = 0, tile_i = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_i) {
        const int num_kb = p.phase_tap_count[tile / tiles_per_phase] * p.cin_blocks;
        const uint32_t acc = tile_i & 1;
        const uint32_t acc_ph = (tile_i >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * Cfg::kAccCols;
        for (int kb = 0; kb < num_kb; kb += Cfg::kKB) {
          const int nk = min(Cfg::kKB, num_kb - kb);
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t st = smem_base + s * Cfg::kStageBytes;
          if (elect_one()) {
            if (dbg != 2) {
#pragma unroll
              for (int j = 0; j < Cfg::kKB; ++j) {
                if (j < nk) {
                  const uint32_t sa = st + j * Cfg::kKbBytes;
                  const uint64_t da = make_desc_kmajor(sa);
                  const uint64_t db = B_MN ? make_desc_mnmajor(sa + Cfg::kABytes, 64) : make_desc_kmajor(sa + Cfg::kABytes);
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    // K advance of 16 elements: +32 B inside the swizzle row (K-major),
                    // +16 rows * 128 B (MN-major); descriptor address field is in 16-byte units
                    const uint64_t da_k = da + static_cast<uint64_t>(k * 2);
                    const uint64_t db_k = db + static_cast<uint64_t>(B_MN ? k * 128 : k * 2);
                    umma_bf16(tmem_d, da_k, db_k, idesc, (kb | j | k) != 0 ? 1u : 0u);
                  }
                }
              }
            }
            umma_commit(&empty_bar[s]);        // frees the smem slot when these MMAs retire
          }
          __syncwarp();
          if (++s == static_cast<uint32_t>(nstages)) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(&tmem_full_bar[acc]);    // accumulator complete
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ============================== epilogue ==============================
    const int e = threadIdx.x - 128;          // 0..127 == tile row == TMEM lane
    const int ew = warp - 4;                  // TMEM lane quarter
    uint32_t tile_i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_i) {
      const int phase = tile / tiles_per_phase;
      const int tp = tile - phase * tiles_per_phase;
      const int mt = tp / n_tiles;
      const int n0 = (tp - mt * n_tiles) * BLOCK_N;
      const int img = mt / tiles_per_img;
      const int t_in = mt - img * tiles_per_img;
      const int th = t_in / p.tiles_w;
      const int w0 = (t_in - th * p.tiles_w) * p.bw;
      const int h0 = th * p.bh;
      const uint32_t acc = tile_i & 1;
      const uint32_t acc_ph = (tile_i >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * Cfg::kAccCols + (static_cast<uint32_t>(ew * 32) << 16);

      if (dbg == 1) { tc_fence_before(); mbar_arrive(&tmem_empty_bar[acc]); continue; }
      conv_epilogue_tile<BLOCK_N>(p, p.out + p.phase_out_off[phase], stage_buf, red, taddr, img, h0, w0, n0, e, ew, lane, true,
                                  [&]() { mbar_arrive(&tmem_empty_bar[acc]); });
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
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
  const int total = p.tiles_w * p.tiles_h * num_images * ((p.cout + BLOCK_N - 1) / BLOCK_N) * p.num_phases;
  const int dbg_stages = g_opt.stages, dbg_grid = g_opt.grid;
  int nst = Cfg::kStages;
  if (dbg_stages > 0 && dbg_stages < nst) nst = dbg_stages;
  int cap = num_sms();
  if (dbg_grid > 0 && dbg_grid < cap) cap = dbg_grid;
  const int grid = total < cap ? total : cap;
  conv_gemm_kernel<BLOCK_N, B_MN><<<grid, 256, Cfg::kSmemBytes, stream>>>(p, num_images, nst, dbg_flag());
}

bool launch_conv_gemm_2cta(const ConvGemmParams& p, int block_n, bool b_mn_major, int num_images,
                           cudaStream_t stream);
bool launch_conv_gemm_ts(const ConvGemmParams& p, int block_n, bool b_mn_major, int num_images,
                         cudaStream_t stream);

int dbg_flag() { return g_opt.dbg; }
bool use_2cta() { return g_opt.two_cta; }
void set_kernel_options(int dbg, bool two_cta, int stages, int grid) {
  g_opt.dbg = dbg;
  g_opt.two_cta = two_cta;
  g_opt.stages = stages;
  g_opt.grid = grid;
}

void launch_conv_gemm(const ConvGemmParams& p, int block_n, bool b_mn_major, int num_images,
                      cudaStream_t stream) {
  // wide layers: CTA pairs (cta_group::2) halve the weight-tile traffic per SM
  if (p.fp8) {
    if (!launch_conv_gemm_2cta(p, block_n, b_mn_major, num_images, stream)) {
      fprintf(stderr, "launch_conv_gemm: fp8 needs block_n == 256, K-major B\n");
      abort();
    }
    return;
  }
  if (p.ts.enabled && use_2cta() && p.tiles_w * p.tiles_h * num_images >= 2 &&
      launch_conv_gemm_ts(p, block_n, b_mn_major, num_images, stream))
    return;
  if (block_n >= 128 && use_2cta() && p.tiles_w * p.tiles_h * num_images >= 2 &&
      launch_conv_gemm_2cta(p, block_n, b_mn_major, num_images, stream))
    return;
  if (!b_mn_major) {
    switch (block_n) {
      case 16: return launch_t<16, false>(p, num_images, stream);
      case 32: return launch_t<32, false>(p, num_images, stream);
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
