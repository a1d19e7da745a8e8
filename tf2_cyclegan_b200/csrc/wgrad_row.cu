// Tap-shared, 2-CTA weight-gradient GEMM for 3x3 stride-1 reflect-padded convs (see
// wgrad_gemm.cuh). Motivation (measured): the generic wgrad kernel is bound by L2 -> SM
// operand traffic (~7 TB/s), not by the tensor pipe; sharing one activation tile between the
// three horizontal taps and one dY tile between all of them halves the bytes per FLOP.
// The shifted reads rely on SWIZZLE_128B being a function of the absolute shared-memory
// address (verified by tools/shift_probe.py on B200: a descriptor may start at any 128-B row).
#include <stdlib.h>

#include "wgrad_gemm.cuh"

namespace b2 {

int dbg_flag();

namespace {

constexpr int kDyBytes = 2 * 64 * 128;        // 128 co x 64 px
constexpr int kActBytes = 9 * 1024;           // 66 px x 128 B = 8448 B, padded to a 1 KB multiple
constexpr int kStageBytes = kDyBytes + kActBytes;
constexpr int kStages = 7;
constexpr int kTmemCols = 512;                // 3 accumulators x 128 columns
constexpr int kRowStage = 128 * 4 + 16;       // epilogue staging row (128 fp32 + pad)
constexpr int kBarrierBytes = (2 * kStages + 1) * 8 + 16;
constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + 1024;

__global__ void __launch_bounds__(256, 1)
wgrad_row_kernel(const __grid_constant__ WgradRowParams p, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_tiles = p.cin / 128;
  const int pair = blockIdx.x >> 1;
  const int m0 = (pair / n_tiles) * 256 + static_cast<int>(rank) * 128;   // this CTA's 128 output channels
  const int n0 = (pair % n_tiles) * 128;                                   // the pair's 128 input channels
  const int r = blockIdx.y;                                                // kernel row
  const int kt_per_img = p.tiles_w * p.H;
  const int total_kt = kt_per_img * p.num_images;
  const int per = (total_kt + gridDim.z - 1) / gridDim.z;
  const int kt_begin = blockIdx.z * per;
  const int kt_end = min(total_kt, kt_begin + per);
  const int num_kb = kt_end - kt_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmDY);
    tma_prefetch_desc(&p.tmA);
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 2);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 2) tmem_alloc_2sm(tmem_ptr_smem, kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (num_kb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int kb = 0; kb < num_kb; ++kb) {
          const int kt = kt_begin + kb;
          const int img = kt / kt_per_img;
          const int t_in = kt - img * kt_per_img;
          const int oh = t_in / p.tiles_w;
          const int w0 = (t_in - oh * p.tiles_w) * 64;
          const int s = kb % kStages;
          const uint32_t ph = (kb / kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * kStageBytes;
          uint8_t* sb = sa + kDyBytes;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * (kDyBytes + 66 * 128));
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_4d_2sm(sa + j * (64 * 128), &p.tmDY, &full_bar[s], m0 + j * 64, w0, oh, img);
          // padded activation row oh + r, pixels [w0, w0 + 66): covers the three taps s = 0, 1, 2
          tma_load_4d_2sm(sb, &p.tmA, &full_bar[s], n0 + static_cast<int>(rank) * 64, w0, oh + r, img);
          if (rank != 0) mbar_arrive_cluster(&full_bar[s], 0);
        }
      }
    } else if (warp == 1) {
      if (lane == 0 && rank == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(256, 128, 1, 1);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int s = kb % kStages;
          const uint32_t ph = (kb / kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * kStageBytes);
          const uint32_t sb = sa + kDyBytes;
          const uint64_t da = make_desc_mnmajor(sa, 64);
          const uint64_t db = make_desc_mnmajor(sb, 64);
          if (dbg != 2) {
#pragma unroll
            for (int tap = 0; tap < 3; ++tap) {
#pragma unroll
              for (int k = 0; k < 4; ++k)   // tap shift: +1 pixel row = +128 B = +8 descriptor units
                umma_bf16_2sm(tmem_base + tap * 128, da + static_cast<uint64_t>(k * 128),
                              db + static_cast<uint64_t>(tap * 8 + k * 128), idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit_2sm(&empty_bar[s]);
        }
        umma_commit_2sm(tmem_full_bar);
      }
    } else if (warp >= 4) {
      const int e = threadIdx.x - 128;
      const int ew = warp - 4;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      const int co = m0 + e;
      float* srow = reinterpret_cast<float*>(smem + e * kRowStage);
      if (dbg != 1) {
        for (int tap = 0; tap < 3; ++tap) {
          const uint32_t taddr = tmem_base + tap * 128 + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
          for (int c = 0; c < 128; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(taddr + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<uint4*>(srow + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          fence_proxy_async_smem();
          if (co < p.cout) {
            float* grow = p.dw + static_cast<long long>(co) * p.ldw + (r * 3 + tap) * p.cin + n0;
            bulk_reduce_add_f32(grow, srow, 128 * 4);
            tma_store_commit();
            tma_store_wait_all();          // the row buffer is reused by the next tap
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, kTmemCols);
}

}  // namespace

void launch_wgrad_row(const WgradRowParams& p, int ksplit, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(wgrad_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * (p.cout / 256) * (p.cin / 128), 3, ksplit);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, wgrad_row_kernel, p, dbg_flag());
}

}  // namespace b2
