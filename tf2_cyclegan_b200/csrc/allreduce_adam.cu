// Two-shot all-reduce fused with the optimizer, executed from inside ONE kernel
// with peer-to-peer (or NVLS multimem) loads and stores over NVLink/NVSwitch:
//
//   (push form) stage_q[me][..] = my grad of shard q, for every peer q: posted peer stores
//   barrier(peers)                      all ranks' gradients are final (pull) / have landed (push)
//   for i in my shard:                  rank r owns elements [r*n/W, (r+1)*n/W)
//     g  = sum_p grad_p[i]              pull: P2P loads | multimem.ld_reduce; push: local staging reads
//     w,m,v = keras_adam(w,m,v,g)       fp32 master + moments live only on the owner (ZeRO-1)
//     w_bf16_p[i] = bf16(w) for all p   P2P stores  | multimem.st (switch broadcast)
//   (one-shot part, small vectors)      every rank reduces + updates its replica, no broadcast
//   (scalars)                           every rank sums all ranks' loss slots
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

// peer loads: volatile (kept after the barrier, in program order) but without a memory clobber, so
// the compiler may keep several of them — and the local master / moment loads — in flight together
__device__ __forceinline__ float4 ld_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 mc_ld_reduce_f4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void mc_st_bf16x8(__nv_bfloat16* p, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(p),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void acc4(float4& a, const float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// Keras-Adam, 4 elements. sqrt / divide use the SFU approximations (2 MUFU + a few FMAs per element
// instead of ~25 instructions for the IEEE forms): with a handful of CTAs the fused kernel is otherwise
// bound by this arithmetic, not by NVLink or HBM. Relative error ~2^-22, far below the bf16 weights.
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void adam4(float4& w, float4& m, float4& v, const float4 g, float lr_t,
                                      float b1, float b2, float eps) {
#define B2_ADAM1(c)                                  \
  m.c = b1 * m.c + (1.f - b1) * g.c;                 \
  v.c = b2 * v.c + (1.f - b2) * g.c * g.c;           \
  w.c = w.c - __fdividef(lr_t * m.c, fast_sqrt(v.c) + eps);
  B2_ADAM1(x) B2_ADAM1(y) B2_ADAM1(z) B2_ADAM1(w)
#undef B2_ADAM1
}

// Keras-Adam on 8 consecutive elements of the local fp32 master / moments; returns the packed bf16 weights
__device__ __forceinline__ uint4 adam8(const AllreduceAdamParams& p, long long i, const float4 g0, const float4 g1,
                                       float lr_t) {
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
  return pk;
}

// A CTA of the collective owns its SM anyway (the conv GEMMs use the whole register file, so nothing
// co-resides with it), therefore it uses ALL of it — pull form: 512 threads x up to 128 registers (many
// peer loads in flight per thread); push form: 1024 threads x 64 registers — and both forms are launched
// as cluster pairs (one TPC per pair), so that g CTAs displace g / 2 of the GEMMs' CTA pairs, not g.
constexpr int kPullThreads = 512;
constexpr int kPushThreads = 1024;
constexpr int kPushUnroll = 4;  // push phase: chunks per thread per iteration (local loads in flight)

__device__ __forceinline__ void st_peer_f4(float* p, const float4 v) {
  asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w) : "memory");
}

// MC: the two-shot reduction goes through the NVSwitch multicast alias (pull form only). An in-switch
// reduction over 8 sources has a long round trip (~20 us measured), so that path keeps 4 chunk pairs
// = 8 multimem.ld_reduce.v4 per thread in flight; the P2P code (two peers x two chunks of loads) is
// compiled out of it, which is what makes the registers available.
template <bool PUSH, bool MC>
__global__ void __launch_bounds__(PUSH ? kPushThreads : kPullThreads, 1)
allreduce_adam_kernel(const AllreduceAdamParams p) {
  constexpr int kArThreads = PUSH ? kPushThreads : kPullThreads;
  constexpr int kArUnroll = PUSH ? 1 : (MC ? 4 : 2);   // 8-element chunks per thread per iteration
  const int t = *p.step + 1;
  const float lr_t = p.lr * sqrtf(1.f - powf(p.beta2, static_cast<float>(t))) /
                     (1.f - powf(p.beta1, static_cast<float>(t)));
  // two-shot: shard of this rank, in units of 8 elements, and the slice of this block inside it.
  // Block b of every rank works on sub-slice b of every shard, so the per-block pairwise barriers
  // below order exactly the data that block exchanges.
  const long long n8 = p.n / 8;
  const long long per_rank = (n8 + p.world - 1) / p.world;
  const long long r_begin = min(n8, per_rank * p.rank), r_end = min(n8, r_begin + per_rank);
  const long long per_blk = (per_rank + gridDim.x - 1) / gridDim.x;
  const long long b_begin = min(r_end, r_begin + per_blk * blockIdx.x);
  const long long b_end = min(r_end, b_begin + per_blk);
  float* gl = p.grad[p.rank];
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);

  if constexpr (PUSH) {
    // ---- push: my gradients of shard q, sub-slice b -> rank q's staging slot [me]; then zero them.
    // Safe without an opening barrier: the closing barrier of the PREVIOUS launch on this lane (all
    // cross-GPU launches share one lane and one global order) proves every peer has finished
    // reading its staging buffer.
    for (int dq = 1; dq < p.world; ++dq) {
      const int q = (p.rank + dq) % p.world;          // staggered: at any time every rank targets another peer
      const long long q_begin = min(n8, per_rank * q), q_end = min(n8, q_begin + per_rank);
      const long long s0 = min(q_end, q_begin + per_blk * blockIdx.x), s1 = min(q_end, s0 + per_blk);
      float* dst = p.stage[q] + (static_cast<long long>(p.rank) * per_rank - q_begin) * 8;   // + i8 * 8
      for (long long i8 = s0 + threadIdx.x; i8 < s1; i8 += kArThreads * kPushUnroll) {
        float4 a[kPushUnroll], b[kPushUnroll];
#pragma unroll
        for (int j = 0; j < kPushUnroll; ++j) {
          const long long c8 = i8 + j * kArThreads;
          if (c8 < s1) {
            a[j] = *reinterpret_cast<const float4*>(gl + p.lo + c8 * 8);
            b[j] = *reinterpret_cast<const float4*>(gl + p.lo + c8 * 8 + 4);
          }
        }
#pragma unroll
        for (int j = 0; j < kPushUnroll; ++j) {
          const long long c8 = i8 + j * kArThreads;
          if (c8 < s1) {
            st_peer_f4(dst + c8 * 8, a[j]);
            st_peer_f4(dst + c8 * 8 + 4, b[j]);
            *reinterpret_cast<float4*>(gl + p.lo + c8 * 8) = z;
            *reinterpret_cast<float4*>(gl + p.lo + c8 * 8 + 4) = z;
          }
        }
      }
    }
  }

  block_barrier(p, 0);

  // ---- scalars (block 0): every rank sums all ranks' slots in rank order
  if (blockIdx.x == 0 && threadIdx.x < p.sc_n) {
    float acc = 0.f;
    for (int q = 0; q < p.world; ++q) {
      float v;
      asm volatile("ld.global.relaxed.sys.f32 %0, [%1];" : "=f"(v) : "l"(p.sc_src[q] + threadIdx.x));
      acc += v;
    }
    p.sc_dst[threadIdx.x] = acc;
  }

  // ---- two-shot region: reduce my shard, Adam, broadcast the bf16 weights
  for (long long i8 = b_begin + threadIdx.x; i8 < b_end; i8 += kArThreads * kArUnroll) {
    float4 g0[kArUnroll], g1[kArUnroll];
    long long idx[kArUnroll];
#pragma unroll
    for (int j = 0; j < kArUnroll; ++j) {
      const long long c8 = i8 + j * kArThreads;
      idx[j] = c8 < b_end ? p.lo + c8 * 8 : -1;
    }
    if constexpr (PUSH) {
      // local reads only: my own gradients + the world - 1 staging slots the peers filled
      const float* st = p.stage[p.rank] - r_begin * 8 - p.lo;      // + slot * per_rank * 8 + idx
#pragma unroll
      for (int j = 0; j < kArUnroll; ++j)
        if (idx[j] >= 0) {
          g0[j] = *reinterpret_cast<const float4*>(gl + idx[j]);
          g1[j] = *reinterpret_cast<const float4*>(gl + idx[j] + 4);
          *reinterpret_cast<float4*>(gl + idx[j]) = z;
          *reinterpret_cast<float4*>(gl + idx[j] + 4) = z;
        }
      for (int q = 0; q < p.world; ++q) {             // rank order: the sum does not depend on who owns it
        if (q == p.rank) continue;
        const float* sq = st + static_cast<long long>(q) * per_rank * 8;
#pragma unroll
        for (int j = 0; j < kArUnroll; ++j)
          if (idx[j] >= 0) {
            acc4(g0[j], *reinterpret_cast<const float4*>(sq + idx[j]));
            acc4(g1[j], *reinterpret_cast<const float4*>(sq + idx[j] + 4));
          }
      }
    } else if constexpr (MC) {
#pragma unroll
      for (int j = 0; j < kArUnroll; ++j)
        if (idx[j] >= 0) {
          g0[j] = mc_ld_reduce_f4(p.mc_grad + idx[j]);
          g1[j] = mc_ld_reduce_f4(p.mc_grad + idx[j] + 4);
        }
    } else {
#pragma unroll
      for (int j = 0; j < kArUnroll; ++j) g0[j] = g1[j] = z;
      for (int q0 = 0; q0 < p.world; q0 += 2) {     // two peers' loads in flight together
        float4 a[2][kArUnroll], b[2][kArUnroll];
        const int qa = (p.rank + q0) % p.world;     // staggered so that peers are hit evenly
        const int qb = (p.rank + q0 + 1) % p.world;
        const bool two = q0 + 1 < p.world;
#pragma unroll
        for (int j = 0; j < kArUnroll; ++j)
          if (idx[j] >= 0) {
            a[0][j] = ld_f4(p.grad[qa] + idx[j]);
            b[0][j] = ld_f4(p.grad[qa] + idx[j] + 4);
            if (two) {
              a[1][j] = ld_f4(p.grad[qb] + idx[j]);
              b[1][j] = ld_f4(p.grad[qb] + idx[j] + 4);
            }
          }
#pragma unroll
        for (int j = 0; j < kArUnroll; ++j)
          if (idx[j] >= 0) {
            acc4(g0[j], a[0][j]); acc4(g1[j], b[0][j]);
            if (two) { acc4(g0[j], a[1][j]); acc4(g1[j], b[1][j]); }
          }
      }
    }
#pragma unroll
    for (int j = 0; j < kArUnroll; ++j) {
      if (idx[j] < 0) continue;
      const uint4 pk = adam8(p, idx[j], g0[j], g1[j], lr_t);
      if (p.mc_w_bf16 != nullptr) {
        mc_st_bf16x8(p.mc_w_bf16 + idx[j], pk);
      } else {
#pragma unroll 8
        for (int q = 0; q < p.world; ++q) {
          const int peer = (p.rank + q) % p.world;
          *reinterpret_cast<uint4*>(p.w_bf16[peer] + idx[j]) = pk;
        }
      }
    }
  }

  // ---- one-shot region: replicated update, bit-identical on every rank (fixed peer order 0..W-1)
  const long long v8 = p.v_n / 8;
  for (long long i8 = static_cast<long long>(blockIdx.x) * kArThreads + threadIdx.x; i8 < v8;
       i8 += static_cast<long long>(gridDim.x) * kArThreads) {
    const long long i = p.v_lo + i8 * 8;
    float4 g0 = z, g1 = z;
    for (int q = 0; q < p.world; q += 2) {
      const bool two = q + 1 < p.world;
      const float4 a0 = ld_f4(p.grad[q] + i), b0 = ld_f4(p.grad[q] + i + 4);
      float4 a1 = z, b1 = z;
      if (two) { a1 = ld_f4(p.grad[q + 1] + i); b1 = ld_f4(p.grad[q + 1] + i + 4); }
      acc4(g0, a0); acc4(g1, b0);     // ((g + r0) + r1) ...: the same order on every rank
      acc4(g0, a1); acc4(g1, b1);
    }
    const uint4 pk = adam8(p, i, g0, g1, lr_t);
    *reinterpret_cast<uint4*>(p.w_bf16[p.rank] + i) = pk;
  }

  block_barrier(p, 1);

  if constexpr (!PUSH) {
    // zero the slices (peer, block) of the local gradient buffer: every peer's block with this
    // index has finished reading them (it passed the second barrier)
    for (int q = 0; q < p.world; ++q) {
      const long long q_begin = min(n8, per_rank * q), q_end = min(n8, q_begin + per_rank);
      const long long s0 = min(q_end, q_begin + per_blk * blockIdx.x), s1 = min(q_end, s0 + per_blk);
      for (long long i8 = s0 + threadIdx.x; i8 < s1; i8 += kArThreads) {
        *reinterpret_cast<float4*>(gl + p.lo + i8 * 8) = z;
        *reinterpret_cast<float4*>(gl + p.lo + i8 * 8 + 4) = z;
      }
    }
  }
  for (long long i8 = static_cast<long long>(blockIdx.x) * kArThreads + threadIdx.x; i8 < v8;
       i8 += static_cast<long long>(gridDim.x) * kArThreads) {
    *reinterpret_cast<float4*>(gl + p.v_lo + i8 * 8) = z;
    *reinterpret_cast<float4*>(gl + p.v_lo + i8 * 8 + 4) = z;
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
  const bool push = p.stage[0] != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_blocks);
  cfg.blockDim = dim3(push ? kPushThreads : kPullThreads);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (num_blocks % 2 == 0) ? 2 : 1;   // CTA pairs share a TPC
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (push) cudaLaunchKernelEx(&cfg, allreduce_adam_kernel<true, false>, p);
  else if (p.mc_grad != nullptr) cudaLaunchKernelEx(&cfg, allreduce_adam_kernel<false, true>, p);
  else cudaLaunchKernelEx(&cfg, allreduce_adam_kernel<false, false>, p);
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
