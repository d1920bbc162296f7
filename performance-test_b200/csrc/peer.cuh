// Device-side primitives of the peer-memory path (structures and protocol: peer.h).
#pragma once
#include "peer.h"
#include "sync_ops.cuh"
#include <cuda_runtime.h>

namespace ptb
{

/// LL record of two doubles: four 8-byte words, each = 32 payload bits | epoch << 32, so that every
/// aligned 8-byte store carries its own flag (no fence between data and flag). One thread.
__device__ __forceinline__ void ll_write(unsigned long long* dst, unsigned int epoch, double v0, double v1)
{
  const unsigned long long b0 = static_cast<unsigned long long>(__double_as_longlong(v0));
  const unsigned long long b1 = static_cast<unsigned long long>(__double_as_longlong(v1));
  const unsigned long long tag = static_cast<unsigned long long>(epoch) << 32;
  st_relaxed_sys(dst + 0, (b0 & 0xffffffffull) | tag);
  st_relaxed_sys(dst + 1, (b0 >> 32) | tag);
  st_relaxed_sys(dst + 2, (b1 & 0xffffffffull) | tag);
  st_relaxed_sys(dst + 3, (b1 >> 32) | tag);
}
/// The same record with only its first NW words in use (NW = 1: epoch only, a plain arrival;
/// NW = 2: one double; NW = 4: two doubles) -- fewer L2 requests when 148 CTAs poll 148 records.
template <int NW>
__device__ __forceinline__ void ll_write_n(unsigned long long* dst, unsigned int epoch, double v0, double v1)
{
  const unsigned long long b0 = static_cast<unsigned long long>(__double_as_longlong(v0));
  const unsigned long long b1 = static_cast<unsigned long long>(__double_as_longlong(v1));
  const unsigned long long tag = static_cast<unsigned long long>(epoch) << 32;
  st_relaxed_sys(dst + 0, (NW > 1 ? (b0 & 0xffffffffull) : 0ull) | tag);
  if constexpr (NW > 1)
    st_relaxed_sys(dst + 1, (b0 >> 32) | tag);
  if constexpr (NW > 2)
  {
    st_relaxed_sys(dst + 2, (b1 & 0xffffffffull) | tag);
    st_relaxed_sys(dst + 3, (b1 >> 32) | tag);
  }
}
template <int NW>
__device__ __forceinline__ bool ll_try_read_n(const unsigned long long* src, unsigned int epoch, double& v0, double& v1)
{
  unsigned long long w[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
  for (int i = 0; i < NW; ++i)
    w[i] = ld_relaxed_sys(src + i);
  bool ok = true;
#pragma unroll
  for (int i = 0; i < NW; ++i)
    ok = ok && static_cast<unsigned int>(w[i] >> 32) == epoch;
  v0 = __longlong_as_double(static_cast<long long>((w[0] & 0xffffffffull) | (w[1] << 32)));
  v1 = __longlong_as_double(static_cast<long long>((w[2] & 0xffffffffull) | (w[3] << 32)));
  return ok;
}
/// One attempt: true (and the values) when all four words carry `epoch`.
__device__ __forceinline__ bool ll_try_read(const unsigned long long* src, unsigned int epoch, double& v0, double& v1)
{
  unsigned long long w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w[i] = ld_relaxed_sys(src + i);
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    ok = ok && static_cast<unsigned int>(w[i] >> 32) == epoch;
  v0 = __longlong_as_double(static_cast<long long>((w[0] & 0xffffffffull) | (w[1] << 32)));
  v1 = __longlong_as_double(static_cast<long long>((w[2] & 0xffffffffull) | (w[3] << 32)));
  return ok;
}
/// Spin until all four words of the record carry `epoch`, then decode.
__device__ __forceinline__ void ll_read(const unsigned long long* src, unsigned int epoch, double& v0, double& v1)
{
  unsigned long long w[4];
  bool ok;
  do
  {
    ok = true;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
      w[i] = ld_relaxed_sys(src + i);
      ok = ok && static_cast<unsigned int>(w[i] >> 32) == epoch;
    }
  } while (!ok);
  v0 = __longlong_as_double(static_cast<long long>((w[0] & 0xffffffffull) | (w[1] << 32)));
  v1 = __longlong_as_double(static_cast<long long>((w[2] & 0xffffffffull) | (w[3] << 32)));
}

/// Publish this rank's two partial sums for reduction `epoch` to every rank (one thread).
__device__ __forceinline__ void peer_publish(const PeerView& P, unsigned int epoch, double v0,
                                             double v1)
{
  const unsigned long long b0 = static_cast<unsigned long long>(__double_as_longlong(v0));
  const unsigned long long b1 = static_cast<unsigned long long>(__double_as_longlong(v1));
  const unsigned long long tag = static_cast<unsigned long long>(epoch) << 32;
  const unsigned long long w[4] = {(b0 & 0xffffffffull) | tag, (b0 >> 32) | tag,
                                   (b1 & 0xffffffffull) | tag, (b1 >> 32) | tag};
  for (int r = 0; r < P.nranks; ++r)
  {
    unsigned long long* dst = P.win[r]->red[epoch & 3u][P.rank];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      st_relaxed_sys(dst + i, w[i]);
  }
}

/// Wait for all ranks' partials of reduction `epoch` and add them in rank order (one thread).
__device__ __forceinline__ void peer_collect(const PeerView& P, unsigned int epoch, double& s0,
                                             double& s1)
{
  s0 = 0.0, s1 = 0.0;
  const PeerWindow* W = P.win[P.rank];
  for (int r = 0; r < P.nranks; ++r)
  {
    const unsigned long long* src = W->red[epoch & 3u][r];
    unsigned long long w[4];
    bool ok;
    do
    {
      ok = true;
#pragma unroll
      for (int i = 0; i < 4; ++i)
      {
        w[i] = ld_relaxed_sys(src + i);
        ok = ok && static_cast<unsigned int>(w[i] >> 32) == epoch;
      }
    } while (!ok);
    s0 += __longlong_as_double(static_cast<long long>((w[0] & 0xffffffffull) | (w[1] << 32)));
    s1 += __longlong_as_double(static_cast<long long>((w[2] & 0xffffffffull) | (w[3] << 32)));
  }
}

} // namespace ptb
