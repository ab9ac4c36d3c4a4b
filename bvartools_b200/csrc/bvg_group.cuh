// bvg_group.cuh -- "group of threads that owns one chain" abstraction.
//
//   SerialGroup   one thread (chain per thread on the device; also what tests/emu instantiates with g++)
//   TileGroup<G>  G <= 32 lanes of a warp (several chains per warp), sync = __syncwarp(mask)
//   BlockGroup    the whole CTA (one chain per CTA), sync = __syncthreads()
//
// Contract used by all algorithms: every thread of the group executes the same sequence of sync()/sum() calls;
// work inside a phase is distributed with `for (i = g.tid(); i < N; i += g.size())`.
#pragma once
#include "bvg_common.cuh"

namespace bvg {

struct SerialGroup {
  static constexpr bool kCheapBcast = true;
  static constexpr bool kIsBlock = false;
  static constexpr bool kWarp32 = false;
  BVG_HD int tid() const { return 0; }
  BVG_HD int size() const { return 1; }
  BVG_HD void sync() const {}
  BVG_HD double sum(double v) const { return v; }
  BVG_HD double bcast(double v, int) const { return v; }
};

#if defined(__CUDACC__)
template <int G>
struct TileGroup {
  static constexpr bool kCheapBcast = true;   // register shuffles
  static constexpr bool kIsBlock = false;
  static constexpr bool kWarp32 = (G == 32);  // a full warp per chain: DMMA tile paths apply
  unsigned mask;
  int lane;   // 0..G-1
  __device__ __forceinline__ TileGroup() {
    const int wl = threadIdx.x & 31;
    lane = wl % G;
    mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl - lane));
  }
  __device__ __forceinline__ int tid() const { return lane; }
  __device__ __forceinline__ int size() const { return G; }
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
  // butterfly: every lane ends with the bit-identical sum (each level adds the same two partial sums)
  __device__ __forceinline__ double sum(double v) const {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, G);
    return v;
  }
  __device__ __forceinline__ double bcast(double v, int src) const { return __shfl_sync(mask, v, src, G); }
};

struct BlockGroup {
  static constexpr bool kCheapBcast = false;  // two __syncthreads per broadcast
  static constexpr bool kIsBlock = true;      // k x k blocks run on warp 0 alone (TileGroup<32>): no CTA barriers inside
  static constexpr bool kWarp32 = false;
  double* scratch;   // >= 33 doubles of shared memory
  __device__ __forceinline__ explicit BlockGroup(double* s) : scratch(s) {}
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int size() const { return blockDim.x; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ double sum(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();                       // scratch may still be read from a previous call
    if ((threadIdx.x & 31) == 0) scratch[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += scratch[i];   // same order in every thread -> identical result
    return t;
  }
  __device__ __forceinline__ double bcast(double v, int src) const {
    __syncthreads();
    if ((int)threadIdx.x == src) scratch[32] = v;
    __syncthreads();
    return scratch[32];
  }
};
#endif

}  // namespace bvg
