// Fused gradient all-reduce + Keras-Adam + weight broadcast over NVLink peer
// memory (the one communication-bound hot path of data-parallel CycleGAN).
#pragma once
#include "common.cuh"

namespace b2 {

constexpr int kMaxRanks = 8;

struct AllreduceAdamParams {
  float* grad[kMaxRanks];            // every rank's fp32 gradient buffer (peer-mapped, symmetric)
  __nv_bfloat16* w_bf16[kMaxRanks];  // every rank's bf16 compute copy of the weights
  uint32_t* signal[kMaxRanks];       // every rank's signal pad: [blocks][world] flags, zero-initialised
  float* w;                          // local fp32 master weights (only this rank's shard is live)
  float* m;
  float* v;
  const int* step;                   // device counter of applied updates
  float lr, beta1, beta2, eps;
  long long n;                       // elements (multiple of 8)
  int rank, world;
  float* mc_grad;                    // NVLS multicast alias of the gradient buffer, or nullptr
  __nv_bfloat16* mc_w_bf16;          // NVLS multicast alias of the bf16 weights, or nullptr
};

void launch_allreduce_adam(const AllreduceAdamParams& p, int num_blocks, cudaStream_t s);
void launch_scatter_add(float* dst, const float* src, const int* idx, long long n, cudaStream_t s);
void launch_gather_bf16(__nv_bfloat16* dst, const float* src, const int* idx, long long n,
                        cudaStream_t s);
void launch_gather_bf16_from_bf16(__nv_bfloat16* dst, const __nv_bfloat16* src, const int* idx, long long n,
                                  cudaStream_t s);

}  // namespace b2
