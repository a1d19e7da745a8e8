// MXFP8 implicit-GEMM convolution for sm_100a: tcgen05.mma.kind::mxf8f6f4.block_scale with E4M3
// activations and weights and one E8M0 scale per 32 K elements of either operand (the OCP MX
// format; BASELINE config 3). K block = 128 channels (128 bytes), four K = 32 MMAs per block.
// Two kernels: `conv_gemm_mx_kernel` (one CTA, M = 128 x N = 128 tiles, one activation box per filter
// tap: the tap-walking TMA scheme of conv_gemm.cu) and, further down, `conv_gemm_mx_ts_kernel` (CTA
// pair, tap-shared halo tiles: the fast path, used whenever the halo tile fits). The first block of
// comments describes what they share and the per-tap kernel.
//
// Scale factors live in TMEM next to the accumulators, in the layout the instruction reads
// (taken from the CUTLASS headers vendored in this image, cute/atom/mma_traits_sm100.hpp
// `tmem_sf_frg`): for operand row r = r0 + 32 r1 the four scales of a 128-byte K block are the
// four bytes of TMEM column r1, lane r0, replicated in each of the four 32-lane sub-partitions;
// the MMA for K slice k selects byte k through the scale-factor id fields of the instruction
// descriptor. Activation scales come from a pixel-major record buffer (16 bytes per pixel, written
// by the InstanceNorm-apply kernel) that the producer loads with the same shifted TMA box as the
// activations; a scale warpgroup transposes them into TMEM with tcgen05.st (TMA cannot deliver
// the interleaved order from a pixel-major tensor). Weight scales are pre-arranged per
// (tile, tap, K block) and fetched with one 512-byte bulk copy.
//
// Per-tap kernel, warp roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM
// allocator, warps 4-7 epilogue, warps 8-11 scale factors. TMEM: columns [0, 256) two accumulators,
// [256, 256 + 8 * stages) scale factors (4 columns SFA + 4 columns SFB per stage).
#include <cuda_fp8.h>
#include <stdlib.h>

#include "conv_epilogue.cuh"
#include "mx.cuh"

namespace b2 {

int dbg_flag();
bool use_2cta();

namespace {

constexpr int kN = 128;
struct MxCfg {
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = kN * 128;
  static constexpr int kSfaBytes = 128 * kMxSfRecord;
  static constexpr int kSfbBytes = kN * 4;
  static constexpr int kTxBytes = kABytes + kBBytes + kSfaBytes + kSfbBytes;
  static constexpr int kStageBytes = (kTxBytes + 1023) / 1024 * 1024;   // A / B tiles need 1024-byte alignment
  static constexpr int kStages = 5;
  static constexpr int kAccCols = kN;
  static constexpr int kSfCol0 = 2 * kAccCols;
  static constexpr int kTmemCols = 512;
  static constexpr int kStagingBytes = EpiCfg<kN>::kStagingBytes;
  static constexpr int kRedBytes = EpiCfg<kN>::kRedBytes;
  static constexpr int kBarrierBytes = (3 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kRedBytes + kBarrierBytes + 1024;
  static_assert(kSfCol0 + 8 * kStages <= kTmemCols, "TMEM budget");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// kind::mxf8f6f4 instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptorBlockScaled):
// [4,6) B scale id, [7,10) A format (0 = E4M3), [10,13) B format, 15 / 16 A / B major (0 = K),
// [17,23) N / 8, 23 scale format (1 = E8M0), [24,29) M / 16, [29,31) A scale id.
__host__ __device__ constexpr uint32_t make_idesc_mxf8(int M, int N) {
  return (uint32_t(N >> 3) << 17) | (1u << 23) | (uint32_t(M >> 4) << 24);
}

__device__ __forceinline__ void umma_mxf8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}

// thread = TMEM lane; writes 4 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b),
               "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(384, 1)
conv_gemm_mx_kernel(const __grid_constant__ ConvGemmMxParams q, int num_images, int dbg) {
  using Cfg = MxCfg;
  const ConvGemmParams& p = q.g;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* stage_buf = smem + Cfg::kStages * Cfg::kStageBytes;
  float* red = reinterpret_cast<float*>(stage_buf + Cfg::kStagingBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_buf + Cfg::kStagingBytes + Cfg::kRedBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* sf_bar = empty_bar + Cfg::kStages;           // scale factors of the stage are in TMEM
  uint64_t* tmem_full_bar = sf_bar + Cfg::kStages;       // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int m_tiles = tiles_per_img * num_images;
  const int n_tiles = (p.cout + kN - 1) / kN;
  const int total_tiles = m_tiles * n_tiles;
  const int kb_per_tile = p.num_taps * p.cin_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&q.tmSFA);
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1 + 4);     // MMA commit + the four scale warps (they read the stage's scale bytes)
      mbar_init(&sf_bar[s], 4);            // one arrival per scale warp
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
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / n_tiles;
      const int nt = tile - mt * n_tiles;
      const int img = mt / tiles_per_img;
      const int t_in = mt - img * tiles_per_img;
      const int th = t_in / p.tiles_w;
      const int w0 = (t_in - th * p.tiles_w) * p.bw;
      const int h0 = th * p.bh;
      const uint8_t* sfb_tile = q.sfb + static_cast<long long>(nt) * kb_per_tile * Cfg::kSfbBytes;
      for (int t = 0; t < p.num_taps; ++t) {
        const Tap tap = p.taps[t];
        for (int cb = 0; cb < p.cin_blocks; ++cb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          uint8_t* sfa = sb + Cfg::kBBytes;
          uint8_t* sfb = sfa + Cfg::kSfaBytes;
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[s], Cfg::kTxBytes);
            tma_load_4d(sa, &p.tmA[0], &full_bar[s], cb * 128, w0 + tap.dw, h0 + tap.dh, img + tap.dn);
            tma_load_2d(sb, &p.tmB, &full_bar[s], tap.bk + cb * 128, nt * kN);             // box {128, 64}:
            tma_load_2d(sb + 64 * 128, &p.tmB, &full_bar[s], tap.bk + cb * 128, nt * kN + 64);   // two halves
            tma_load_4d(sfa, &q.tmSFA, &full_bar[s], 0, w0 + tap.dw, h0 + tap.dh, img + tap.dn);
            const int tk = tap.bk / (p.cin_blocks * 128);   // position of the tap in the weight matrix
            bulk_load(sfb, sfb_tile + static_cast<long long>(cb * p.num_taps + tk) * Cfg::kSfbBytes,
                      Cfg::kSfbBytes, &full_bar[s]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ============================= MMA issuer =============================
    constexpr uint32_t idesc = make_idesc_mxf8(128, kN);
    const uint32_t smem_base = smem_u32(smem);
    uint32_t it = 0, tile_i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_i) {
      const uint32_t acc = tile_i & 1;
      const uint32_t acc_ph = (tile_i >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * Cfg::kAccCols;
      for (int kb = 0; kb < kb_per_tile; ++kb, ++it) {
        const int s = it % Cfg::kStages;
        const uint32_t ph = (it / Cfg::kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        mbar_wait(&sf_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * Cfg::kStageBytes;
        const uint64_t da = make_desc_kmajor(sa);
        const uint64_t db = make_desc_kmajor(sa + Cfg::kABytes);
        const uint32_t tsfa = tmem_base + Cfg::kSfCol0 + 8 * s;
        const uint32_t tsfb = tsfa + 4;
        if (elect_one()) {
          if (dbg != 2) {
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
              // K slice k: +32 bytes inside the swizzle row; scale byte k of the stage's columns
              // (id fields of the descriptor and the two top bits of the scale address)
              umma_mxf8(tmem_d, da + k * 2, db + k * 2, idesc | (k << 29) | (k << 4), tsfa | (k << 30),
                        tsfb | (k << 30), (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&tmem_full_bar[acc]);
      __syncwarp();
    }
  } else if (warp >= 8) {
    // =========================== scale factors ============================
    // thread (sub-partition sp, lane) writes TMEM lane 32 sp + lane: SFA column j = the 4 scale
    // bytes of tile pixel lane + 32 j, SFB column j = those of output channel lane + 32 j
    const int sp = warp - 8;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(sp * 32) << 16) + Cfg::kSfCol0;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int t = 0; t < p.num_taps; ++t) {
        for (int cb = 0; cb < p.cin_blocks; ++cb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          const uint8_t* sfa = smem + s * Cfg::kStageBytes + Cfg::kABytes + Cfg::kBBytes;
          const uint8_t* sfb = sfa + Cfg::kSfaBytes;
          uint32_t a[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            a[j] = *reinterpret_cast<const uint32_t*>(sfa + (lane + 32 * j) * kMxSfRecord + (cb & 3) * 4);
          const uint4 b = *reinterpret_cast<const uint4*>(sfb + lane * 16);
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[s]);   // this warp's shared-memory reads of the stage are done
          if (dbg != 5) {   // timing experiment 5: no TMEM stores
            tmem_st_32x4(lane_base + 8 * s, a[0], a[1], a[2], a[3]);
            tmem_st_32x4(lane_base + 8 * s + 4, b.x, b.y, b.z, b.w);
            tmem_st_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sf_bar[s]);     // one arrival per warp: 128 per stage would serialise on the barrier
        }
      }
    }
  } else if (warp >= 4) {
    // ============================== epilogue ==============================
    const int e = threadIdx.x - 128;
    const int ew = warp - 4;
    uint32_t tile_i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_i) {
      const int mt = tile / n_tiles;
      const int n0 = (tile - mt * n_tiles) * kN;
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
      conv_epilogue_tile<kN>(p, p.out, stage_buf, red, taddr, img, h0, w0, n0, e, ew, lane, true,
                             [&]() { mbar_arrive(&tmem_empty_bar[acc]); });
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------
// Tap-shared CTA-pair variant (the fast path; structure of conv_gemm_ts.cu): M = 256 pixels across
// the pair (cta_group::2), N = 128, each CTA loads half of the weight tile, ONE activation halo
// tile + ONE scale-record halo tile per 128-channel block serve all taps — the MMA reads the
// activations through a row-shifted descriptor and the scale warpgroup reads the scale records
// of the same shifted rows. Operand traffic per 128 x 128 output tile drops from 621 KB to ~200 KB,
// which is what the per-tap kernel above is bound by.
//
// Warp roles (512 threads per CTA): 0 activation / weight producer, 1 MMA issuer (leader CTA), 2 TMEM
// allocator, 3 scale-data producer, 4-7 and 12-15 the two epilogue warpgroups (64 accumulator columns
// each), 8-11 scale warpgroup. Barriers: a_full / b_full / sf_bar / tmem_* live in the leader (both CTAs'
// operand loads complete there, as in conv_gemm_ts.cu); the scale data of each CTA lands on ITS OWN
// sfd_full barrier because its own scale warpgroup consumes it.
// Optional stall attribution (tools/kernel_bench.py --mx_profile): cycles every role of CTA 0 spends in
// each of its barrier waits, written to a 16-slot device buffer. Off (nullptr) in normal runs.
__device__ unsigned long long* g_mx_prof = nullptr;
struct WaitClock {   // adds the cycles of its scope to a LOCAL counter (flushed once at the end of the role)
  long long* acc;
  long long t0;
  __device__ __forceinline__ WaitClock(long long* counters, int i, bool on) : acc(on ? counters + i : nullptr), t0(0) {
    if (acc) t0 = clock64();
  }
  __device__ __forceinline__ ~WaitClock() {
    if (acc) *acc += clock64() - t0;
  }
};

struct MxTsCfg {
  static constexpr int kMaxTaps = 9;                       // 3x3 filters (forward and flat dgrad)
  static constexpr int kAStageBytes = 34 * 1024;           // activation halo tile (<= 272 rows x 128 B)
  static constexpr int kSfHaloBytes = 5 * 1024;            // scale records of the same rows (16 B each)
  static constexpr int kSfbBytes = kN * 4;                 // weight scales of one (tap, 128-channel block)
  static constexpr int kSfbAllBytes = 5 * 1024;            // ... of all taps of the block (<= 9 x 512 B)
  static constexpr int kAStride = kAStageBytes;
  static constexpr int kAStages = 2;
  // scale data (records of the halo pixels + weight scales of all taps) has its own, deeper ring fed by
  // its own producer warp: the scale warpgroup needs ~2000 cycles per block, so it must run more than one
  // block ahead of the MMAs (a block's MMAs take about as long)
  static constexpr int kSfStride = kSfHaloBytes + kSfbAllBytes;
  static constexpr int kSfStages = 3;
  static constexpr int kBBytes = 64 * 128;                 // this CTA's half of the 128-row weight tile
  // Taps per weight stage: the MMA warp's own instruction stream costs ~140 cycles per barrier wait and
  // ~50 per tcgen05.mma (time line: profiles/r2/mx_timeline.txt) against 64 tensor cycles per MMA, so a
  // stage carries three taps (one wait / one commit for twelve MMAs), and ALL scale factors of a
  // 128-channel block (every tap) are staged into TMEM once, off the per-stage critical path.
  static constexpr int kTaps = 3;
  static constexpr int kBStride = kTaps * kBBytes;
  static constexpr int kBStages = 3;
  static constexpr int kSfCols = 8 * kMaxTaps;             // TMEM columns per activation stage (SFA + SFB per tap)
  static constexpr int kAccCols = kN;
  static constexpr int kSfCol0 = 2 * kAccCols;
  static constexpr int kTmemCols = 512;
  // two epilogue warpgroups, 64 columns each: with one, draining a 128 x 128 tile (~7000 cycles) takes
  // longer than its MMAs (~5200) and the issue warp waits for a free accumulator
  static constexpr int kHalfN = kN / 2;
  static constexpr int kStagingBytes = 2 * EpiCfg<kHalfN>::kStagingBytes;
  static constexpr int kRedBytes = 2 * EpiCfg<kHalfN>::kRedBytes;
  static constexpr int kNumBars = 2 * kAStages + 3 * kSfStages + 2 * kBStages + 4;
  static constexpr int kBarrierBytes = kNumBars * 8 + 16;
  static constexpr int kSmemBytes = kAStages * kAStride + kSfStages * kSfStride + kBStages * kBStride + kStagingBytes +
                                    kRedBytes + kBarrierBytes + 1024;
  static_assert(kAStride % 1024 == 0 && kBStride % 1024 == 0 && kSfStride % 1024 == 0 &&
                kMaxTaps * kSfbBytes <= kSfbAllBytes, "stage layout");
  static_assert(kSfCol0 + kSfCols * kSfStages <= kTmemCols, "TMEM budget");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

__device__ __forceinline__ void umma_mxf8_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}

template <bool PROF>
__global__ void __launch_bounds__(512, 1)
conv_gemm_mx_ts_kernel(const __grid_constant__ ConvGemmMxParams q, int num_images, int dbg) {
  using Cfg = MxTsCfg;
  const ConvGemmParams& p = q.g;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + Cfg::kAStages * Cfg::kAStride;
  uint8_t* sf_ring = b_ring + Cfg::kBStages * Cfg::kBStride;
  uint8_t* stage_buf = sf_ring + Cfg::kSfStages * Cfg::kSfStride;
  float* red = reinterpret_cast<float*>(stage_buf + Cfg::kStagingBytes);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(stage_buf + Cfg::kStagingBytes + Cfg::kRedBytes);
  uint64_t* a_empty = a_full + Cfg::kAStages;
  uint64_t* sfd_full = a_empty + Cfg::kAStages;          // local: this CTA's scale records + weight scales landed
  uint64_t* sf_empty = sfd_full + Cfg::kSfStages;        // multicast commit: the block's MMAs retired
  uint64_t* sf_bar = sf_empty + Cfg::kSfStages;          // leader: both CTAs' scales of the block are in TMEM
  uint64_t* b_full = sf_bar + Cfg::kSfStages;
  uint64_t* b_empty = b_full + Cfg::kBStages;
  uint64_t* tmem_full_bar = b_empty + Cfg::kBStages;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int m_tiles = tiles_per_img * num_images;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int n_tiles = p.cout / kN;
  const int total_pairs = m_pairs * n_tiles;
  const int num_taps = p.num_taps;
  const TapShare ts = p.ts;
  // first halo row of every tap, and the K offset of its weights
  __shared__ uint32_t tap_row[Cfg::kMaxTaps];
  __shared__ int tap_bk[Cfg::kMaxTaps];
  __shared__ uint32_t tap_col[Cfg::kMaxTaps];   // first scale column of the tap: 8 x its position in the weight matrix
  if (threadIdx.x < num_taps) {
    const Tap tap = p.taps[threadIdx.x];
    tap_bk[threadIdx.x] = tap.bk;
    tap_col[threadIdx.x] = 8u * static_cast<uint32_t>(tap.bk / (p.cin_blocks * 128));
    tap_row[threadIdx.x] = static_cast<uint32_t>((tap.dh - ts.origin_dh) * ts.pitch + (tap.dw - ts.origin_dw));
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmTS);
    tma_prefetch_desc(&q.tmSFTS);
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kAStages; ++s) {
      mbar_init(&a_full[s], 2);            // one arrival per CTA of the pair
      mbar_init(&a_empty[s], 1);           // multicast commit from the leader's MMA thread
    }
    for (int s = 0; s < Cfg::kSfStages; ++s) {
      mbar_init(&sfd_full[s], 1);
      mbar_init(&sf_empty[s], 1 + 4);      // MMA commit + this CTA's four scale warps (readers of the slot)
      mbar_init(&sf_bar[s], 8);            // 4 scale warps x 2 CTAs
    }
    for (int s = 0; s < Cfg::kBStages; ++s) {
      mbar_init(&b_full[s], 2);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 512);  // 256 epilogue threads x 2 CTAs
    }
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 2) tmem_alloc_2sm(tmem_ptr_smem, Cfg::kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  auto decode = [&](int pt, int& mt, int& nt, int& img, int& h0, int& w0) {
    const int mp = pt / n_tiles;
    nt = pt - mp * n_tiles;
    mt = 2 * mp + static_cast<int>(rank);
    img = mt / tiles_per_img;              // == num_images for the phantom tile: TMA zero-fills
    const int t_in = mt - img * tiles_per_img;
    const int th = t_in / p.tiles_w;
    w0 = (t_in - th * p.tiles_w) * p.bw;
    h0 = th * p.bh;
  };
  const uint32_t sf_box_bytes = static_cast<uint32_t>(ts.box_bytes) / 128u * kMxSfRecord;
  // stall attribution (PROF only): per-role cycle counters flushed at the end + a time line of the MMA warp
  unsigned long long* prof_out = g_mx_prof;
  const bool prof_on = PROF && prof_out != nullptr && blockIdx.x == 0 && lane == 0 &&
                       (warp == 0 || warp == 1 || warp == 8 || warp == 4);
  long long prof[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) prof[i] = 0;
  const long long role_t0 = prof_on ? clock64() : 0;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    uint32_t ita = 0, itb = 0;
    for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
      int mt, nt, img, h0, w0;
      decode(pt, mt, nt, img, h0, w0);
      for (int cb = 0; cb < p.cin_blocks; ++cb, ++ita) {
        const int sa = ita % Cfg::kAStages;
        { WaitClock wc(prof, 1, prof_on); mbar_wait(&a_empty[sa], ((ita / Cfg::kAStages) & 1) ^ 1); }
        uint8_t* abuf = a_ring + sa * Cfg::kAStride;
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&a_full[sa], 2 * ts.nbox * ts.box_bytes);
          for (int b = 0; b < ts.nbox; ++b)
            tma_load_4d_2sm(abuf + b * ts.box_bytes, &p.tmTS, &a_full[sa], cb * 128,
                            w0 + ts.origin_dw + b * ts.box_step, h0 + ts.origin_dh, img);
          if (rank != 0) mbar_arrive_cluster(&a_full[sa], 0);
        }
        __syncwarp();
        for (int t = 0; t < num_taps; t += Cfg::kTaps, ++itb) {
          const int nk = min(Cfg::kTaps, num_taps - t);
          const int sb = itb % Cfg::kBStages;
          { WaitClock wc(prof, 2, prof_on); mbar_wait(&b_empty[sb], ((itb / Cfg::kBStages) & 1) ^ 1); }
          uint8_t* bbuf = b_ring + sb * Cfg::kBStride;
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(&b_full[sb], 2 * nk * Cfg::kBBytes);
            for (int j = 0; j < nk; ++j)
              tma_load_2d_2sm(bbuf + j * Cfg::kBBytes, &p.tmB, &b_full[sb], tap_bk[t + j] + cb * 128,
                              nt * kN + rank * 64);
            if (rank != 0) mbar_arrive_cluster(&b_full[sb], 0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 3) {
    // ===================== scale-data producer (both CTAs, own barriers) =====================
    uint32_t its = 0;
    for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
      int mt, nt, img, h0, w0;
      decode(pt, mt, nt, img, h0, w0);
      for (int cb = 0; cb < p.cin_blocks; ++cb, ++its) {
        const int ss = its % Cfg::kSfStages;
        mbar_wait(&sf_empty[ss], ((its / Cfg::kSfStages) & 1) ^ 1);
        uint8_t* sbuf = sf_ring + ss * Cfg::kSfStride;
        if (elect_one()) {
          mbar_arrive_expect_tx(&sfd_full[ss], ts.nbox * sf_box_bytes + num_taps * Cfg::kSfbBytes);
          for (int b = 0; b < ts.nbox; ++b)
            tma_load_4d(sbuf + b * sf_box_bytes, &q.tmSFTS, &sfd_full[ss], 0, w0 + ts.origin_dw + b * ts.box_step,
                        h0 + ts.origin_dh, img);
          // weight scales of every tap of this (tile, channel block): contiguous, in weight-matrix tap order
          bulk_load(sbuf + Cfg::kSfHaloBytes,
                    q.sfb + (static_cast<long long>(nt) * p.cin_blocks + cb) * num_taps * Cfg::kSfbBytes,
                    num_taps * Cfg::kSfbBytes, &sfd_full[ss]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_mxf8(256, kN);
      const uint32_t a_ring_u = smem_u32(a_ring), b_ring_u = smem_u32(b_ring);
      uint32_t ita = 0, itb = 0, tile_i = 0;
      // results of non-blocking probes of the NEXT stage's / block's barriers, issued before the current
      // stage's MMAs so that the ~140-cycle barrier latency is off the issue path
      bool b_ready = false, a_ready = false, sf_ready = false;
      // per-tap descriptor offset (halo row x 128 B, in the descriptor's 16-byte units) and scale columns,
      // kept in registers: the issue loop below is fully unrolled, so there is no load -> address -> R2UR
      // chain in front of every tap's first MMA
      uint32_t t_enc[Cfg::kMaxTaps], t_col[Cfg::kMaxTaps];
#pragma unroll
      for (int i = 0; i < Cfg::kMaxTaps; ++i) {
        t_enc[i] = i < num_taps ? tap_row[i] * 8u : 0u;
        t_col[i] = i < num_taps ? tap_col[i] : 0u;
      }
      // high word of the activation descriptor (SWIZZLE_128B, version 1, stride between 8-row groups)
      const uint32_t da_hi = ((static_cast<uint32_t>(ts.sbo) >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      constexpr uint32_t db_hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters, ++tile_i) {
        const uint32_t acc = tile_i & 1;
        { WaitClock wc(prof, 4, prof_on); mbar_wait(&tmem_empty_bar[acc], ((tile_i >> 1) & 1) ^ 1); }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * Cfg::kAccCols;
        uint32_t accumulate = 0;
        for (int cb = 0; cb < p.cin_blocks; ++cb, ++ita) {
          const int sa = ita % Cfg::kAStages;
          const uint32_t pha = (ita / Cfg::kAStages) & 1;
          const bool stamp_a = prof_on && ita < 14;   // block-boundary time line (slots 216 ..)
          if (stamp_a) prof_out[216 + 4 * ita + 0] = clock64() - role_t0;
          if (!a_ready) { WaitClock wc(prof, 5, prof_on); mbar_wait(&a_full[sa], pha); }
          if (stamp_a) prof_out[216 + 4 * ita + 1] = clock64() - role_t0;
          const int ss = ita % Cfg::kSfStages;
          if (!sf_ready) { WaitClock wc(prof, 7, prof_on); mbar_wait(&sf_bar[ss], (ita / Cfg::kSfStages) & 1); }
          if (stamp_a) prof_out[216 + 4 * ita + 2] = clock64() - role_t0;
          tc_fence_after();
          const uint32_t a_enc = ((a_ring_u + sa * Cfg::kAStride) >> 4) & 0x3FFFu;
          const uint32_t tsf_a = tmem_base + Cfg::kSfCol0 + Cfg::kSfCols * ss;
#pragma unroll
          for (int st = 0; st < Cfg::kMaxTaps / Cfg::kTaps; ++st) {
            const int t = st * Cfg::kTaps;
            if (t >= num_taps) break;
            const int nk = min(Cfg::kTaps, num_taps - t);
            const int sb = itb % Cfg::kBStages;
            const bool stamp = prof_on && itb < 40;     // time line of the first 40 stages (slots 16 ..)
            if (stamp) prof_out[16 + 5 * itb + 0] = clock64() - role_t0;
            if (!b_ready) { WaitClock wc(prof, 6, prof_on); mbar_wait(&b_full[sb], (itb / Cfg::kBStages) & 1); }
            if (stamp) prof_out[16 + 5 * itb + 1] = clock64() - role_t0;
            tc_fence_after();
            {   // probe what the next iteration(s) will wait for
              const uint32_t nb = itb + 1;
              b_ready = mbar_test_wait(&b_full[nb % Cfg::kBStages], (nb / Cfg::kBStages) & 1);
              if (t + Cfg::kTaps >= num_taps) {
                const uint32_t na = ita + 1;
                a_ready = mbar_test_wait(&a_full[na % Cfg::kAStages], (na / Cfg::kAStages) & 1);
                sf_ready = mbar_test_wait(&sf_bar[na % Cfg::kSfStages], (na / Cfg::kSfStages) & 1);
              }
            }
            const uint32_t b_enc = ((b_ring_u + sb * Cfg::kBStride) >> 4) & 0x3FFFu;
            if (elect_one()) {
              if (dbg != 2) {
                WaitClock wc(prof, 13, prof_on);
#pragma unroll
                for (int j = 0; j < Cfg::kTaps; ++j) {
                  if (j < nk) {
                    const uint64_t da = (static_cast<uint64_t>(da_hi) << 32) | (a_enc + t_enc[t + j]);
                    const uint64_t db = (static_cast<uint64_t>(db_hi) << 32) | (b_enc + j * (Cfg::kBBytes >> 4));
                    const uint32_t tsfa = tsf_a + t_col[t + j], tsfb = tsfa + 4;
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k)   // K slice k: scale byte k, selected by the descriptor's id fields
                      umma_mxf8_2sm(tmem_d, da + k * 2, db + k * 2, idesc | (k << 29) | (k << 4), tsfa, tsfb,
                                    (j == 0 && k == 0) ? accumulate : 1u);
                  }
                }
              }
              if (stamp) prof_out[16 + 5 * itb + 3] = clock64() - role_t0;
              {
                WaitClock wc(prof, 14, prof_on);
                umma_commit_2sm(&b_empty[sb]);
              }
            }
            __syncwarp();
            if (stamp) prof_out[16 + 5 * itb + 4] = clock64() - role_t0;
            accumulate = 1;
            ++itb;
          }
          if (elect_one()) {
            umma_commit_2sm(&a_empty[sa]);
            umma_commit_2sm(&sf_empty[ss]);
          }
          __syncwarp();
          if (stamp_a) prof_out[216 + 4 * ita + 3] = clock64() - role_t0;
        }
        if (elect_one()) umma_commit_2sm(&tmem_full_bar[acc]);
        __syncwarp();
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // ===================== scale factors (both CTAs, own TMEM) =====================
    // once per 128-channel block: the scales of EVERY tap go to TMEM (8 columns per tap: 4 SFA + 4 SFB),
    // so nothing of this is on the MMA warp's per-stage path
    const int sp = warp - 8;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(sp * 32) << 16) + Cfg::kSfCol0;
    const uint32_t group_rows = static_cast<uint32_t>(ts.sbo) / 128u;   // halo rows between 8-row groups of the tile
    uint32_t rel[4];                        // halo row of tile pixel lane + 32 i, relative to the tap's first row
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t m = lane + 32 * i;
      rel[i] = (m >> 3) * group_rows + (m & 7);
    }
    // per-tap halo byte offset / scale column in registers (static indexing in the unrolled loops below)
    uint32_t t_off[Cfg::kMaxTaps], t_col[Cfg::kMaxTaps];
#pragma unroll
    for (int i = 0; i < Cfg::kMaxTaps; ++i) {
      t_off[i] = i < num_taps ? tap_row[i] * kMxSfRecord : 0u;
      t_col[i] = i < num_taps ? tap_col[i] : 0u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) rel[i] *= kMxSfRecord;
    uint32_t ita = 0;
    for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
      for (int cb = 0; cb < p.cin_blocks; ++cb, ++ita) {
        const int ss = ita % Cfg::kSfStages;
        { WaitClock wc(prof, 9, prof_on); mbar_wait(&sfd_full[ss], (ita / Cfg::kSfStages) & 1); }
        const uint8_t* sbuf = sf_ring + ss * Cfg::kSfStride;
        const uint8_t* sfh = sbuf + (cb & 3) * 4;
        const uint8_t* sfb = sbuf + Cfg::kSfHaloBytes + lane * 16;
        const uint32_t col0 = lane_base + Cfg::kSfCols * ss;
        // all loads first (independent, their latencies overlap), then the TMEM stores
        uint32_t a[Cfg::kMaxTaps][4];
        uint4 b[Cfg::kMaxTaps];
#pragma unroll
        for (int t = 0; t < Cfg::kMaxTaps; ++t) {
          if (t < num_taps) {
#pragma unroll
            for (int i = 0; i < 4; ++i) a[t][i] = *reinterpret_cast<const uint32_t*>(sfh + t_off[t] + rel[i]);
            b[t] = *reinterpret_cast<const uint4*>(sfb + t_col[t] * (Cfg::kSfbBytes / 8));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sf_empty[ss]);   // this warp's shared-memory reads of the slot are done
        if (dbg != 5) {   // timing experiment 5: no TMEM stores
#pragma unroll
          for (int t = 0; t < Cfg::kMaxTaps; ++t) {
            if (t < num_taps) {
              tmem_st_32x4(col0 + t_col[t], a[t][0], a[t][1], a[t][2], a[t][3]);
              tmem_st_32x4(col0 + t_col[t] + 4, b[t].x, b[t].y, b[t].z, b[t].w);
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&sf_bar[ss], 0);   // one arrival per warp
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own TMEM lanes) =====================
    // warps 4-7 and 12-15: one warpgroup per 64-column half of the tile
    const int wg = warp >= 12 ? 1 : 0;
    const int e = threadIdx.x & 127;
    const int ew = warp & 3;
    uint8_t* my_stage = stage_buf + wg * EpiCfg<Cfg::kHalfN>::kStagingBytes;
    float* my_red = red + wg * (EpiCfg<Cfg::kHalfN>::kRedBytes / 4);
    uint32_t tile_i = 0;
    for (int pt = cluster_id; pt < total_pairs; pt += num_clusters, ++tile_i) {
      int mt, nt, img, h0, w0;
      decode(pt, mt, nt, img, h0, w0);
      const uint32_t acc = tile_i & 1;
      { WaitClock wc(prof, 12, prof_on && warp == 4); mbar_wait(&tmem_full_bar[acc], (tile_i >> 1) & 1); }
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * Cfg::kAccCols + wg * Cfg::kHalfN + (static_cast<uint32_t>(ew * 32) << 16);
      if (dbg == 1) { tc_fence_before(); mbar_arrive_cluster(&tmem_empty_bar[acc], 0); continue; }
      conv_epilogue_tile<Cfg::kHalfN>(p, p.out, my_stage, my_red, taddr, img, h0, w0, nt * kN + wg * Cfg::kHalfN, e, ew,
                                      lane, mt < m_tiles, [&]() { mbar_arrive_cluster(&tmem_empty_bar[acc], 0); }, dbg,
                                      1 + 2 * wg);
    }
  }

  if (prof_on) {   // each role flushes its own counter slots
    const int lo = warp == 0 ? 0 : warp == 1 ? 3 : warp == 8 ? 8 : 11;
    const int hi = warp == 0 ? 3 : warp == 1 ? 8 : warp == 8 ? 11 : 13;
    prof[lo] = clock64() - role_t0;
    for (int i = lo; i < hi; ++i) prof_out[i] = static_cast<unsigned long long>(prof[i]);
    if (warp == 1) for (int i = 13; i < 16; ++i) prof_out[i] = static_cast<unsigned long long>(prof[i]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
}

// one thread per (row, 32-element K block)
__global__ void __launch_bounds__(256) mx_weight_quant_kernel(const MxWeightQuantParams p) {
  const int per_tap = p.kdim / kMxBlock;
  const int kblocks = p.taps * per_tap;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.rows * kblocks) return;
  const int n = idx / kblocks, kb = idx - n * kblocks;
  const int tap = kb / per_tap, r = kb - tap * per_tap;
  const __nv_bfloat16* src = p.w + n * p.s_row + tap * p.s_tap + static_cast<long long>(r) * kMxBlock * p.s_k;
  float v[32];
  float amax = 0.f;
  if (p.s_k == 1) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = __ldg(s4 + i);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        v[i * 8 + 2 * j] = f.x;
        v[i * 8 + 2 * j + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __bfloat162float(src[i * p.s_k]);
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) amax = fmaxf(amax, fabsf(v[i]));
  const uint32_t sb = mx_scale_byte(amax);
  const float inv = mx_inv_scale(sb);
  uint32_t o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(v[4 * j] * inv, v[4 * j + 1] * inv), __NV_SATFINITE, __NV_E4M3);
    const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(v[4 * j + 2] * inv, v[4 * j + 3] * inv), __NV_SATFINITE, __NV_E4M3);
    o[j] = lo | (hi << 16);
  }
  uint4* dst = reinterpret_cast<uint4*>(p.w8 + (static_cast<long long>(n) * kblocks + kb) * kMxBlock);
  dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
  dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
  // scale position: [n / 128][K block of 128][tap][n % 32][(n % 128) / 32][K slice]
  const int k_blocks = p.kdim / 128;
  const long long chunk = (static_cast<long long>(n / 128) * k_blocks + (r >> 2)) * p.taps + tap;
  p.sfb[chunk * 512 + (n & 31) * 16 + ((n & 127) >> 5) * 4 + (r & 3)] = static_cast<uint8_t>(sb);
}

}  // namespace

// device buffer of 16 counters (or 0 to switch the stall attribution off)
static bool g_mx_prof_host = false;
void set_mx_profile(unsigned long long* buf) {
  cudaMemcpyToSymbol(g_mx_prof, &buf, sizeof(buf));
  g_mx_prof_host = buf != nullptr;     // selects the instrumented instantiation of the kernel
}

void launch_conv_gemm_mx(const ConvGemmMxParams& p, int num_images, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(conv_gemm_mx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MxCfg::kSmemBytes);
    configured = true;
  }
  if (p.g.num_phases != 1 || p.g.cout % kN != 0 || p.g.cin_blocks < 1 || p.g.cin_blocks > 4) {
    fprintf(stderr, "launch_conv_gemm_mx: needs one phase, cout %% 128 == 0, cin in {128, ..., 512}\n");
    abort();
  }
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const TapShare& ts = p.g.ts;
  if (ts.enabled && use_2cta() && p.g.tiles_w * p.g.tiles_h * num_images >= 2 && ts.nbox >= 1 &&
      p.g.num_taps <= MxTsCfg::kMaxTaps &&
      ts.nbox * ts.box_bytes <= MxTsCfg::kAStageBytes && (ts.nbox == 1 || ts.box_bytes % 1024 == 0)) {
    static bool configured_ts = false;
    if (!configured_ts) {
      cudaFuncSetAttribute(conv_gemm_mx_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MxTsCfg::kSmemBytes);
      cudaFuncSetAttribute(conv_gemm_mx_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MxTsCfg::kSmemBytes);
      configured_ts = true;
    }
    const int m_tiles = p.g.tiles_w * p.g.tiles_h * num_images;
    const int pairs = ((m_tiles + 1) / 2) * (p.g.cout / kN);
    const int clusters = pairs < sms / 2 ? pairs : sms / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = MxTsCfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (g_mx_prof_host) cudaLaunchKernelEx(&cfg, conv_gemm_mx_ts_kernel<true>, p, num_images, dbg_flag());
    else cudaLaunchKernelEx(&cfg, conv_gemm_mx_ts_kernel<false>, p, num_images, dbg_flag());
    return;
  }
  const int total = p.g.tiles_w * p.g.tiles_h * num_images * (p.g.cout / kN);
  const int grid = total < sms ? total : sms;
  conv_gemm_mx_kernel<<<grid, 384, MxCfg::kSmemBytes, stream>>>(p, num_images, dbg_flag());
}

void launch_mx_weight_quant(const MxWeightQuantParams& p, cudaStream_t stream) {
  if (p.kdim % 128 != 0 || p.rows % 128 != 0) {
    fprintf(stderr, "launch_mx_weight_quant: rows and K per tap must be multiples of 128\n");
    abort();
  }
  const int total = p.rows * (p.taps * p.kdim / kMxBlock);
  mx_weight_quant_kernel<<<(total + 255) / 256, 256, 0, stream>>>(p);
}

}  // namespace b2
