// Multi-tap wgrad for narrow layers (Cin-side tile of 64: generator stem / output conv, PatchGAN
// conv0): ONE CTA accumulates ALL taps of its K range, so the dY tile is loaded once per K tile
// instead of once per tap (these layers are TMA-issue bound, not tensor bound). T <= 7 taps,
// T accumulators of 64 TMEM columns each. Same operand conventions as wgrad_gemm.cu.
#include <stdlib.h>

#include "wgrad_gemm.cuh"

namespace b2 {

int dbg_flag();

namespace {

constexpr int kBoxBytes = 64 * 128;           // 64 pixels x 64 channels
constexpr int kMaxT = 7;

__device__ __forceinline__ void red_row(float* grow, float* srow) {
  bulk_reduce_add_f32(grow, srow, 64 * 4);
  tma_store_commit();
  tma_store_wait_all();
}

__global__ void __launch_bounds__(256, 1)
wgrad_multi_kernel(const __grid_constant__ WgradParams p, int T, int stages) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int stage_bytes = (2 + T) * kBoxBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int ktiles_per_img = p.ktiles_w * p.ktiles_h;
  const int total_kt = ktiles_per_img * p.num_images;
  const int per = (total_kt + gridDim.z - 1) / gridDim.z;
  const int kt_begin = blockIdx.z * per;
  const int kt_end = min(total_kt, kt_begin + per);
  const int num_kb = kt_end - kt_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmDY[0]);
    tma_prefetch_desc(&p.tmA[0]);
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (num_kb > 0) {
    // producer / MMA warps: warp-uniform loops, one elected lane issues
    if (warp == 0) {
      int img = kt_begin / ktiles_per_img;
      int t_in = kt_begin - img * ktiles_per_img;
      int th = t_in / p.ktiles_w;
      int tw = t_in - th * p.ktiles_w;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int w0 = tw * p.kbw, h0 = th * p.kbh;
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * stage_bytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_4d(sa + j * kBoxBytes, &p.tmDY[0], &full_bar[s], m0 + j * 64, w0, h0, img);
          for (int t = 0; t < T; ++t) {
            const Tap tap = p.taps[t];
            tma_load_4d(sa + (2 + t) * kBoxBytes, &p.tmA[tap.map], &full_bar[s], 0, w0 + tap.dw,
                        h0 + tap.dh, img);
          }
        }
        __syncwarp();
        if (++tw == p.ktiles_w) { tw = 0; if (++th == p.ktiles_h) { th = 0; ++img; } }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      const uint32_t smem_base = smem_u32(smem);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * stage_bytes;
        const uint64_t da = make_desc_mnmajor(sa, 64);
        if (elect_one()) {
          for (int t = 0; t < T; ++t) {
            const uint64_t db = make_desc_mnmajor(sa + (2 + t) * kBoxBytes, 64);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + t * 64, da + static_cast<uint64_t>(k * 128),
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
      const int co = m0 + e;
      float* srow = reinterpret_cast<float*>(smem + e * (64 * 4 + 16));
      for (int t = 0; t < T; ++t) {
        const uint32_t taddr = tmem_base + t * 64 + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(srow + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        fence_proxy_async_smem();
        if (co < p.cout) red_row(p.dw + static_cast<long long>(co) * p.ldw + p.taps[t].bk, srow);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// requirements: cin == 64 (one 64-wide activation window per tap), one dY map, <= 7 taps
bool launch_wgrad_multi(const WgradParams& p, int ksplit, cudaStream_t stream) {
  if (p.cin != 64 || p.num_taps > kMaxT || p.num_taps < 2) return false;
  const int T = p.num_taps;
  const int stage_bytes = (2 + T) * kBoxBytes;
  int stages = (225 * 1024 - 1024 - 256) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return false;
  const int smem_bytes = stages * stage_bytes + (2 * stages + 1) * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(wgrad_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    configured = true;
  }
  dim3 grid((p.cout + 127) / 128, 1, ksplit);
  wgrad_multi_kernel<<<grid, 256, smem_bytes, stream>>>(p, T, stages);
  return true;
}

}  // namespace b2
