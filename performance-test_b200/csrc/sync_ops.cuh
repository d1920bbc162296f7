// Scoped loads/stores of the synchronisation protocols (grid barrier, fused halo, peer windows)
// in one place. Device build: the PTX below. PTB_HOST_EMU (tests/emu runs the kernel sources on
// the host, test harness only): the GCC atomics with the same ordering.
#pragma once
#include <cuda_runtime.h>
#ifdef PTB_HOST_EMU
#include <sched.h>
#endif

namespace ptb
{

#ifdef PTB_HOST_EMU
// every scoped load sits in a spin loop somewhere: give the core away so that the hundreds of
// host threads of the harness make progress
template <typename T>
__device__ __forceinline__ T emu_ld(const T* p, int order)
{
  const T v = __atomic_load_n(p, order);
  sched_yield();
  return v;
}
template <typename T>
__device__ __forceinline__ void emu_st(T* p, T v, int order)
{
  __atomic_store_n(p, v, order);
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) { emu_st(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) { return emu_ld(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) { emu_st(p, v, __ATOMIC_RELAXED); }
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) { return emu_ld(p, __ATOMIC_RELAXED); }
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) { return emu_ld(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) { emu_st(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) { return emu_ld(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void st_release_gpu_u32(unsigned int* p, unsigned int v) { emu_st(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ double ld_ca_f64(const double* q) { return *reinterpret_cast<const volatile double*>(q); }
__device__ __forceinline__ void fence_acq_rel_gpu() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
__device__ __forceinline__ unsigned long long l2_policy(int) { return 0ull; }
__device__ __forceinline__ double ld_stream(const double* q, unsigned long long) { return *q; }
__device__ __forceinline__ int ld_stream(const int* q, unsigned long long) { return *q; }
#else
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p)
{
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int* p, unsigned int v)
{
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// coherent cached load: L1 lines are dropped by the acquire of the grid barrier (never the .nc path)
__device__ __forceinline__ double ld_ca_f64(const double* q)
{
  double v;
  asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(q));
  return v;
}
// fence.acq_rel.gpu: orders this thread's (and, through bar.sync, its CTA's) earlier accesses before
// later ones at GPU scope and drops the L1 (CCTL.IVALL, B300_MICROARCH.md) -- what a grid barrier
// needs, without the sequential-consistency round of __threadfence() (MEMBAR.SC).
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// L2 residency control for the operator stream ("L2 plan"). The matrix is read
// once per CG iteration and never again before ~0.5-4 GB of other traffic has passed, while the
// vectors (and a fixed prefix of the matrix, sized to what is left of the 126 MB L2) are re-read
// every iteration: matrix loads carry an L2 eviction-priority hint and bypass L1.
//   kind 0: evict_normal   1: evict_first (streamed part)   2: evict_last (pinned prefix)
__device__ __forceinline__ unsigned long long l2_policy(int kind)
{
  unsigned long long pol;
  if (kind == 1)
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else if (kind == 2)
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else
    asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double ld_stream(const double* q, unsigned long long pol)
{
  double v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(q), "l"(pol));
  return v;
}
__device__ __forceinline__ int ld_stream(const int* q, unsigned long long pol)
{
  int v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(q), "l"(pol));
  return v;
}
#endif

} // namespace ptb
