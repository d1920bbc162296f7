// TEST INFRASTRUCTURE -- runs the source of the persistent CG loop kernel (csrc/cg.cu cg_loop) on
// the host. One loaded copy of this library plays ONE CTA: `__shared__` variables become statics
// of that copy, its 256 threads are std::threads, __syncthreads and the warp collectives are
// barriers. The test loads G copies (separate files -> separate statics) and calls them from G
// host threads at once; they meet in the kernel's own grid barrier, in memory they all share.
// Nothing here is linked into the product libraries; the product path never sees PTB_HOST_EMU.
#define PTB_HOST_EMU 1
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#undef __shared__
#define __shared__ static
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

struct EmuIdx
{
  unsigned x = 0, y = 0, z = 0;
};
static thread_local EmuIdx threadIdx, blockIdx;
static EmuIdx blockDim, gridDim;
static std::barrier<>* emu_cta = nullptr;
static std::barrier<>* emu_warp[32];
static double emu_xd[32][32];
static long long emu_xi[32][32];

static inline void __syncthreads() { emu_cta->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp[threadIdx.x >> 5]->arrive_and_wait(); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
template <typename T>
static inline T __ldcg(const T* p) { return *reinterpret_cast<const volatile T*>(p); }
static inline double2 __ldcg(const double2* p)
{
  const volatile double* q = reinterpret_cast<const volatile double*>(p);
  return double2{q[0], q[1]};
}
template <typename T>
static inline T __ldcv(const T* p) { return *reinterpret_cast<const volatile T*>(p); }
static inline double __shfl_xor_sync(unsigned, double v, int o)
{
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  emu_xd[w][l] = v;
  emu_warp[w]->arrive_and_wait();
  const double r = emu_xd[w][l ^ o];
  emu_warp[w]->arrive_and_wait();
  return r;
}
static inline int __shfl_sync(unsigned, int v, int src)
{
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  emu_xi[w][l] = v;
  emu_warp[w]->arrive_and_wait();
  const int r = static_cast<int>(emu_xi[w][src & 31]);
  emu_warp[w]->arrive_and_wait();
  return r;
}
static inline unsigned __ballot_sync(unsigned, bool pred)
{
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  emu_xi[w][l] = pred ? 1 : 0;
  emu_warp[w]->arrive_and_wait();
  unsigned m = 0;
  for (int i = 0; i < 32; ++i)
    m |= emu_xi[w][i] ? 1u << i : 0u;
  emu_warp[w]->arrive_and_wait();
  return m;
}
static inline bool __all_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline long long __double_as_longlong(double v)
{
  long long r;
  std::memcpy(&r, &v, 8);
  return r;
}
static inline double __longlong_as_double(long long v)
{
  double r;
  std::memcpy(&r, &v, 8);
  return r;
}
using std::max;
using std::min;

#include "../../performance-test_b200/csrc/cg.cu"

extern "C" {

// One CTA (`block` of `grid`) of cg_loop<BS, false> on a single rank. All pointers are shared by
// the G concurrently running copies.
int emu_cg_loop_block(int bs, int block, int grid, int32_t n_rows, int32_t n_slices,
                      const int64_t* mat_off, const int32_t* cols, const double* vals,
                      const int32_t* cdelta, const int32_t* colsx, const int64_t* xoff,
                      const int32_t* order, const double* dinv, double* r, double* p, double* x,
                      double* y, void* st, unsigned long long* slots, int n_it,
                      const int32_t* ounit, const int32_t* bal_begin, int res_cap)
{
  using namespace ptb;
  LoopArgs L{};
  L.A = SpmvArgs{n_rows, n_slices, mat_off, cols, vals, cdelta, colsx, xoff, 0, 0, ounit, bal_begin};
  L.n = static_cast<std::int64_t>(n_rows) * bs;
  L.dinv = dinv, L.r = r, L.p = p, L.x = x, L.y = y;
  L.st = static_cast<CgState*>(st);
  L.slots = slots;
  L.it0 = 0, L.n_it = n_it, L.ebase = 1, L.lbase = 1;
  L.res_cap = bal_begin != nullptr ? res_cap : 0; // x and r of the CTA's run resident in "shared memory"
  PeerView P{};
  P.rank = 0, P.nranks = 1;
  FusedHalo FH{};
  FH.order = order, FH.n_interior = n_slices;
  constexpr unsigned T = SPMV_THREADS;
  gridDim.x = grid, blockDim.x = T;
  std::barrier<> cta(T);
  emu_cta = &cta;
  std::vector<std::unique_ptr<std::barrier<>>> wb;
  for (unsigned w = 0; w < T / 32; ++w)
  {
    wb.push_back(std::make_unique<std::barrier<>>(32));
    emu_warp[w] = wb.back().get();
  }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < T; ++t)
    th.emplace_back([=] {
      threadIdx.x = t, blockIdx.x = block;
      if (bal_begin != nullptr)
        bs == 1 ? cg_loop<1, false, true>(L, P, FH) : cg_loop<3, false, true>(L, P, FH);
      else if (bs == 1)
        cg_loop<1, false>(L, P, FH);
      else
        cg_loop<3, false>(L, P, FH);
    });
  for (auto& t : th)
    t.join();
  return 0;
}

// One CTA of cg_loop<BS, true> (fused halo, peer windows) of rank `rank`; windows / peer_p are the
// ranks' PeerWindow blocks and the neighbours' p vectors, all in host memory shared by the copies.
int emu_cg_loop_block_peer(int bs, int block, int grid, int rank, int nranks, void** windows,
                           int n_nbr, const int* nbr_rank, const int* recv_displ,
                           const double** peer_p, const int32_t* remote_indices,
                           const int32_t* src_index, int n_interior, int npull,
                           unsigned long long* ready, int32_t n_rows, int32_t n_slices,
                           const int64_t* mat_off, const int32_t* cols, const double* vals,
                           const int32_t* cdelta, const int32_t* colsx, const int64_t* xoff,
                           const int32_t* order, const double* dinv, double* r, double* p,
                           double* x, double* y, void* st, unsigned long long* slots, int n_it,
                           const int32_t* ounit, const int32_t* bal_begin, int res_cap)
{
  using namespace ptb;
  LoopArgs L{};
  L.A = SpmvArgs{n_rows, n_slices, mat_off, cols, vals, cdelta, colsx, xoff, 0, 0, ounit, bal_begin};
  L.n = static_cast<std::int64_t>(n_rows) * bs;
  L.dinv = dinv, L.r = r, L.p = p, L.x = x, L.y = y;
  L.st = static_cast<CgState*>(st);
  L.slots = slots;
  L.it0 = 0, L.n_it = n_it, L.ebase = 1, L.lbase = 1;
  L.res_cap = bal_begin != nullptr ? res_cap : 0; // x and r of the CTA's run resident in "shared memory"
  PeerView P{};
  P.rank = rank, P.nranks = nranks;
  for (int q = 0; q < nranks; ++q)
    P.win[q] = static_cast<PeerWindow*>(windows[q]);
  FusedHalo FH{};
  FH.H.n_nbr = n_nbr, FH.H.bs = bs;
  for (int q = 0; q < n_nbr; ++q)
    FH.H.nbr_rank[q] = nbr_rank[q], FH.H.peer_p[q] = peer_p[q];
  for (int q = 0; q <= n_nbr; ++q)
    FH.H.recv_displ[q] = recv_displ[q];
  FH.H.remote_indices = remote_indices, FH.H.src_index = src_index;
  FH.order = order, FH.n_interior = n_interior, FH.npull = npull, FH.epoch = 0;
  FH.ready = ready, FH.pw = p;
  constexpr unsigned T = SPMV_THREADS;
  gridDim.x = grid, blockDim.x = T;
  std::barrier<> cta(T);
  emu_cta = &cta;
  std::vector<std::unique_ptr<std::barrier<>>> wb;
  for (unsigned w = 0; w < T / 32; ++w)
  {
    wb.push_back(std::make_unique<std::barrier<>>(32));
    emu_warp[w] = wb.back().get();
  }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < T; ++t)
    th.emplace_back([=] {
      threadIdx.x = t, blockIdx.x = block;
      if (bal_begin != nullptr)
        bs == 1 ? cg_loop<1, true, true>(L, P, FH) : cg_loop<3, true, true>(L, P, FH);
      else if (bs == 1)
        cg_loop<1, true>(L, P, FH);
      else
        cg_loop<3, true>(L, P, FH);
    });
  for (auto& t : th)
    t.join();
  return 0;
}

int emu_cgstate_size() { return static_cast<int>(sizeof(ptb::CgState)); }
int emu_peerwindow_size() { return static_cast<int>(sizeof(ptb::PeerWindow)); }
}
