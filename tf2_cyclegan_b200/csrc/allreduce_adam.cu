// Two-shot all-reduce fused with the optimizer, executed from inside ONE kernel
// with peer-to-peer (or NVLS multimem) loads and stores over NVLink/NVSwitch:
//
//   barrier(peers)                      all ranks' gradients are final
//   for i in my shard:                  rank r owns elements [r*n/W, (r+1)*n/W)
//     g  = sum_p grad_p[i]              P2P loads   | multimem.ld_reduce (in-switch sum)
//     w,m,v = keras_adam(w,m,v,g)       fp32 master + moments live only on the owner (ZeRO-1)
//     w_bf16_p[i] = bf16(w) for all p   P2P stores  | multimem.st (switch broadcast)
//   barrier(peers)                      everyone's compute copy is complete
//   grad_local[...] = 0                 ready for the next step's wgrad accumulation
//
// No NCCL call, no separate optimizer / cast / memset kernels on this path. The
// kernel uses few CTAs so it can run on a side stream under the next network's
// backward pass. Cross-GPU ordering: CAS flags in each rank's signal pad with
// release / acquire at system scope (self-resetting, so CUDA graphs can replay it).
#include "allreduce_adam.cuh"

namespace b2 {

namespace {

__device__ __forceinline__ void send_signal(uint32_t* addr) {   // remote 0 -> 1, release
  uint32_t old;
  do {
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
  } while (old != 0u);
}
__device__ __forceinline__ void wait_signal(uint32_t* addr) {   // local 1 -> 0, acquire
  uint32_t old;
  long long t0 = clock64();
  do {
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
    if (old != 1u && clock64() - t0 > 8 * B2_WATCHDOG_CYCLES) {
      printf("b2 watchdog: peer signal wait timed out (block %d)\n", blockIdx.x);
      __trap();
    }
  } while (old != 1u);
}

// every block pairs with the same block index on every peer
__device__ __forceinline__ void block_barrier(const AllreduceAdamParams& p, int phase) {
  __syncthreads();
  const int slot_base = (phase * gridDim.x + blockIdx.x) * p.world;
  if (threadIdx.x < p.world) {
    const int peer = threadIdx.x;
    __threadfence_system();
    send_signal(p.signal[peer] + slot_base + p.rank);
    wait_signal(p.signal[p.rank] + slot_base + peer);
  }
  __syncthreads();
}

__device__ __forceinline__ float4 ld_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 mc_ld_reduce_f4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st_bf16x8(__nv_bfloat16* p, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(p),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void adam4(float4& w, float4& m, float4& v, const float4 g, float lr_t,
                                      float b1, float b2, float eps) {
#define B2_ADAM1(c)                                  \
  m.c = b1 * m.c + (1.f - b1) * g.c;                 \
  v.c = b2 * v.c + (1.f - b2) * g.c * g.c;           \
  w.c = w.c - lr_t * m.c / (sqrtf(v.c) + eps);
  B2_ADAM1(x) B2_ADAM1(y) B2_ADAM1(z) B2_ADAM1(w)
#undef B2_ADAM1
}

__global__ void __launch_bounds__(512) allreduce_adam_kernel(const AllreduceAdamParams p) {
  const int t = *p.step + 1;
  const float lr_t = p.lr * sqrtf(1.f - powf(p.beta2, static_cast<float>(t))) /
                     (1.f - powf(p.beta1, static_cast<float>(t)));
  // shard of this rank, in units of 8 elements
  const long long n8 = p.n / 8;
  const long long per_rank = (n8 + p.world - 1) / p.world;
  const long long r_begin = min(n8, per_rank * p.rank), r_end = min(n8, r_begin + per_rank);
  // slice of this block inside the shard
  const long long per_blk = (r_end - r_begin + gridDim.x - 1) / gridDim.x;
  const long long b_begin = min(r_end, r_begin + per_blk * blockIdx.x);
  const long long b_end = min(r_end, b_begin + per_blk);

  block_barrier(p, 0);

  for (long long i8 = b_begin + threadIdx.x; i8 < b_end; i8 += blockDim.x) {
    const long long i = i8 * 8;
    float4 g0, g1;
    if (p.mc_grad != nullptr) {
      g0 = mc_ld_reduce_f4(p.mc_grad + i);
      g1 = mc_ld_reduce_f4(p.mc_grad + i + 4);
    } else {
      g0 = make_float4(0.f, 0.f, 0.f, 0.f);
      g1 = g0;
#pragma unroll 8
      for (int q = 0; q < p.world; ++q) {
        const int peer = (p.rank + q) % p.world;   // stagger so peers are hit evenly
        const float4 a = ld_f4(p.grad[peer] + i);
        const float4 b = ld_f4(p.grad[peer] + i + 4);
        g0.x += a.x; g0.y += a.y; g0.z += a.z; g0.w += a.w;
        g1.x += b.x; g1.y += b.y; g1.z += b.z; g1.w += b.w;
      }
    }
    float4 w0 = *reinterpret_cast<float4*>(p.w + i), w1 = *reinterpret_cast<float4*>(p.w + i + 4);
    float4 m0 = *reinterpret_cast<float4*>(p.m + i), m1 = *reinterpret_cast<float4*>(p.m + i + 4);
    float4 v0 = *reinterpret_cast<float4*>(p.v + i), v1 = *reinterpret_cast<float4*>(p.v + i + 4);
    adam4(w0, m0, v0, g0, lr_t, p.beta1, p.beta2, p.eps);
    adam4(w1, m1, v1, g1, lr_t, p.beta1, p.beta2, p.eps);
    *reinterpret_cast<float4*>(p.w + i) = w0; *reinterpret_cast<float4*>(p.w + i + 4) = w1;
    *reinterpret_cast<float4*>(p.m + i) = m0; *reinterpret_cast<float4*>(p.m + i + 4) = m1;
    *reinterpret_cast<float4*>(p.v + i) = v0; *reinterpret_cast<float4*>(p.v + i + 4) = v1;
    uint4 pk;
    pk.x = pack_bf16x2(w0.x, w0.y); pk.y = pack_bf16x2(w0.z, w0.w);
    pk.z = pack_bf16x2(w1.x, w1.y); pk.w = pack_bf16x2(w1.z, w1.w);
    if (p.mc_w_bf16 != nullptr) {
      mc_st_bf16x8(p.mc_w_bf16 + i, pk);
    } else {
#pragma unroll 8
      for (int q = 0; q < p.world; ++q) {
        const int peer = (p.rank + q) % p.world;
        *reinterpret_cast<uint4*>(p.w_bf16[peer] + i) = pk;
      }
    }
  }

  block_barrier(p, 1);

  // zero the slices (peer, block) of the local gradient buffer: every peer's block with this
  // index has finished reading them (it passed the second barrier)
  float* gl = p.grad[p.rank];
  for (int q = 0; q < p.world; ++q) {
    const long long q_begin = min(n8, per_rank * q), q_end = min(n8, q_begin + per_rank);
    const long long pb = (q_end - q_begin + gridDim.x - 1) / gridDim.x;
    const long long s0 = min(q_end, q_begin + pb * blockIdx.x), s1 = min(q_end, s0 + pb);
    for (long long i8 = s0 + threadIdx.x; i8 < s1; i8 += blockDim.x) {
      *reinterpret_cast<float4*>(gl + i8 * 8) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(gl + i8 * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

__global__ void gather_bf16_kernel(__nv_bfloat16* dst, const float* src, const int* idx, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = idx[i];
    dst[i] = __float2bfloat16(j >= 0 ? src[j] : 0.f);
  }
}

// same gather from the bf16 compute copy (complete on every rank; the fp32 master is sharded)
__global__ void gather_bf16_from_bf16_kernel(__nv_bfloat16* dst, const __nv_bfloat16* src, const int* idx,
                                             long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = idx[i];
    dst[i] = j >= 0 ? src[j] : __float2bfloat16(0.f);
  }
}

__global__ void scatter_add_kernel(float* dst, const float* src, const int* idx, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = idx[i];
    if (j >= 0) atomicAdd(dst + j, src[i]);
  }
}

}  // namespace

void launch_scatter_add(float* dst, const float* src, const int* idx, long long n, cudaStream_t s) {
  long long b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  scatter_add_kernel<<<static_cast<int>(b), 256, 0, s>>>(dst, src, idx, n);
}

void launch_allreduce_adam(const AllreduceAdamParams& p, int num_blocks, cudaStream_t s) {
  allreduce_adam_kernel<<<num_blocks, 512, 0, s>>>(p);
}

void launch_gather_bf16(__nv_bfloat16* dst, const float* src, const int* idx, long long n,
                        cudaStream_t s) {
  long long b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  gather_bf16_kernel<<<static_cast<int>(b), 256, 0, s>>>(dst, src, idx, n);
}

void launch_gather_bf16_from_bf16(__nv_bfloat16* dst, const __nv_bfloat16* src, const int* idx, long long n,
                                  cudaStream_t s) {
  long long b = (n + 255) / 256;
  if (b > 1184) b = 1184;
  if (b < 1) b = 1;
  gather_bf16_from_bf16_kernel<<<static_cast<int>(b), 256, 0, s>>>(dst, src, idx, n);
}

}  // namespace b2
