// Deterministic two-stage reductions (no floating-point atomics).
//
// Every reducing kernel is launched with exactly RED_BLOCKS x RED_THREADS threads and walks
// its data grid-stride, so each thread always sums the same elements in the same order.
// Block partials go to red_part[v][block]; the block that finishes last (integer ticket)
// adds the partials in block order and publishes the result.  Run-to-run bit-reproducible.
#pragma once

#include "fcvm_common.cuh"

namespace fcvm {

template <int NV>
struct Slots {
  int s[NV];
};

template <int NV>
__device__ __forceinline__ void block_reduce_publish(double (&v)[NV], double *red_part, unsigned int *counter,
                                                     double *out, Slots<NV> slots) {
  __shared__ double sm[NV][RED_THREADS / 32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    double w = warp_sum(v[i]);
    if (lane == 0) sm[i][warp] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < RED_THREADS / 32; w++) s += sm[i][w];
      red_part[i * RED_BLOCKS + blockIdx.x] = s;
    }
    __threadfence();
    unsigned int ticket = atomicAdd(counter, 1u);
    last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    // fixed-order sum of the block partials: RED_BLOCKS values per result, one warp each
    for (int i = warp; i < NV; i += RED_THREADS / 32) {
      double s = 0.0;
      for (int b = lane; b < RED_BLOCKS; b += 32) s += __ldcg(&red_part[i * RED_BLOCKS + b]);
      s = warp_sum(s);
      if (lane == 0) out[slots.s[i]] = s;
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

}  // namespace fcvm

namespace fcvm {
template <int NV>
__device__ __forceinline__ void block_reduce_publish(double (&v)[NV], double *red_part, unsigned int *counter,
                                                     double *out) {
  Slots<NV> id;
#pragma unroll
  for (int i = 0; i < NV; i++) id.s[i] = i;
  block_reduce_publish<NV>(v, red_part, counter, out, id);
}
}  // namespace fcvm
