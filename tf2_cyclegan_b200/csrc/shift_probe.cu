// Hardware probe (not on any hot path): can a SWIZZLE_128B K-major operand be read starting
// at a row that is NOT 1024-byte aligned, i.e. can several filter taps share ONE loaded
// activation tile through shifted UMMA descriptors? For every row shift s in 0..7 and for
// base_offset in {0, s} it computes D = A[s : s + 128] * B^T and returns all results.
#include "common.cuh"

namespace b2 {

__global__ void __launch_bounds__(128, 1)
shift_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sa = smem;                 // 136 rows x 128 B
  uint8_t* sb = smem + 18432;         // 16 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 18432 + 2048);
  uint64_t* mma_bar = bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 136 * 128 + 16 * 128);
    tma_load_2d(sa, &tmA, bar, 0, 0);
    tma_load_2d(sb, &tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, 16, 0, 0);
    const uint64_t db = make_desc_kmajor(smem_u32(sb));
    for (int v = 0; v < 16; ++v) {
      const int s = v & 7;
      const int mode = v >> 3;
      uint64_t da = make_desc_kmajor(smem_u32(sa) + s * 128);
      if (mode == 1) da |= static_cast<uint64_t>(s & 7) << 49;   // matrix base offset field
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_base + v * 16, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2),
                  idesc, k != 0 ? 1u : 0u);
    }
    umma_commit(mma_bar);
  }
  mbar_wait(mma_bar, 0);
  tc_fence_after();
  const int row = threadIdx.x;
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  for (int v = 0; v < 16; ++v) {
    uint32_t r[16];
    tmem_ld_32x16(taddr + v * 16, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(v * 128 + row) * 16 + j] = __uint_as_float(r[j]);
  }
  (void)lane;
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

void launch_shift_probe(const CUtensorMap& tmA, const CUtensorMap& tmB, float* out, cudaStream_t s) {
  cudaFuncSetAttribute(shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 24576);
  shift_probe_kernel<<<1, 128, 24576, s>>>(tmA, tmB, out);
}

}  // namespace b2
