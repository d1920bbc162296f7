// Deterministic reductions: fixed shuffle tree inside a warp, fixed order across warps, one partial
// per CTA, and the last CTA to finish sums the partials in index order. No floating-point atomics,
// so a given (grid, block) configuration always produces the same bits.
#pragma once
#include <cuda_runtime.h>

namespace ptb
{

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

/// Sum NV values per thread over the CTA; result valid in thread 0. smem: NV * 32 doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    v[i] = warp_sum(v[i]);
  if (lane == 0)
  {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      smem[i * 32 + warp] = v[i];
  }
  __syncthreads();
  if (warp == 0)
  {
#pragma unroll
    for (int i = 0; i < NV; ++i)
    {
      double t = lane < nwarps ? smem[i * 32 + lane] : 0.0;
      v[i] = warp_sum(t);
    }
  }
}

/// CTA partials -> global sum, finished by the last CTA. `partials` is [NV][gridDim.x].
/// Returns true in every thread of the last CTA, with the totals in out[] (thread 0 only).
/// The ticket counter is reset for the next launch.
template <int NV>
__device__ __forceinline__ bool grid_sum_last_block(double (&v)[NV], double* partials,
                                                    unsigned int* ticket, double* smem,
                                                    double (&out)[NV])
{
  __shared__ bool is_last;
  block_sum<NV>(v, smem);
  if (threadIdx.x == 0)
  {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      partials[i * gridDim.x + blockIdx.x] = v[i];
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last)
    return false;
  __threadfence();
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i)
  {
    acc[i] = 0.0;
    for (unsigned int j = threadIdx.x; j < gridDim.x; j += blockDim.x)
      acc[i] += __ldcg(&partials[i * gridDim.x + j]);
  }
  __syncthreads(); // smem reuse
  block_sum<NV>(acc, smem);
  if (threadIdx.x == 0)
  {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      out[i] = acc[i];
    *ticket = 0u;
  }
  return true;
}

} // namespace ptb
