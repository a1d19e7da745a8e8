// Fused gradient all-reduce + Keras-Adam + weight broadcast over NVLink peer
// memory (the one communication-bound hot path of data-parallel CycleGAN).
#pragma once
#include "common.cuh"

namespace b2 {

constexpr int kMaxRanks = 8;
constexpr int kMaxScalars = 32;

// Two forms of the two-shot part (chosen per launch by the host, see parallel/fused_allreduce.py):
//   pull: barrier | owner loads the peers' shards (P2P ld / multimem.ld_reduce) -> Adam -> broadcast | barrier
//   push: every rank stores the shards it does not own into the owners' staging slots (posted writes)
//         | barrier | owner sums local staging + its own gradients -> Adam -> broadcast | barrier
// One launch covers up to three regions of ONE network's flat buffers (all offsets in elements from
// the buffer bases, multiples of 8):
//   two-shot  [lo, lo + n):      rank r owns shard r (ZeRO-1): reduce peers' gradients -> Adam on the
//                                fp32 master shard -> broadcast bf16 weights to every peer
//   one-shot  [v_lo, v_lo + v_n): small, latency-bound (InstanceNorm gamma / beta, biases): EVERY rank
//                                reduces all of it in the same peer order (bit-identical sums) and runs
//                                Adam on its own replicated fp32 master / moments — no broadcast
//   scalars   sc_n floats:       one-shot SUM of every rank's loss slots into a local result buffer
//                                (replaces the per-step host-issued NCCL reduce of the loss scalars)
// followed, after the closing barrier, by zeroing the local gradients of both regions.
struct AllreduceAdamParams {
  float* grad[kMaxRanks];            // every rank's fp32 gradient buffer (peer-mapped, symmetric)
  __nv_bfloat16* w_bf16[kMaxRanks];  // every rank's bf16 compute copy of the weights
  uint32_t* signal[kMaxRanks];       // every rank's signal pad: [2][blocks][world] flags, zero-initialised
  float* w;                          // local fp32 master weights (two-shot: only this rank's shard is live)
  float* m;
  float* v;
  const int* step;                   // device counter of applied updates
  float lr, beta1, beta2, eps;
  long long lo, n;                   // two-shot region
  long long v_lo, v_n;               // one-shot region (v_n == 0: none)
  const float* sc_src[kMaxRanks];    // every rank's scalar slots (peer-mapped), or nullptr
  float* sc_dst;                     // local [sc_n] sums over ranks
  int sc_n;
  int rank, world;
  float* mc_grad;                    // NVLS multicast alias of the gradient buffer, or nullptr
  __nv_bfloat16* mc_w_bf16;          // NVLS multicast alias of the bf16 weights, or nullptr
  // push mode (stage[0] != nullptr): instead of PULLING the peers' gradients (latency bound: the
  // kernel needs megabytes of loads in flight), every rank PUSHES the shards it does not own into
  // the owner's staging slot with posted peer stores (bandwidth bound with a few CTAs), and the
  // owner reduces local memory. stage[q] = rank q's staging buffer: [world][ceil(n/8/world)*8] floats.
  float* stage[kMaxRanks];
};

void launch_allreduce_adam(const AllreduceAdamParams& p, int num_blocks, cudaStream_t s);
void launch_scatter_add(float* dst, const float* src, const int* idx, long long n, cudaStream_t s);
void launch_gather_bf16(__nv_bfloat16* dst, const float* src, const int* idx, long long n,
                        cudaStream_t s);
void launch_gather_bf16_from_bf16(__nv_bfloat16* dst, const __nv_bfloat16* src, const int* idx, long long n,
                                  cudaStream_t s);

}  // namespace b2
