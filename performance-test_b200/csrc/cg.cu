// ZZZ Solve on the device: linalg::cg of the reference (src/cg.h:38-86) with the optional Jacobi
// extension, fused into three kernels per iteration:
//
//   spmv_sell        y = A p  (the `action` of cg.h:62) + local p.y                 cg.h:62,65
//   cg_update        alpha = rz/py; r -= alpha y; local r.r, r.z                    cg.h:65,71,74
//   cg_direction     x += alpha p; beta = rz'/rz; stopping rule; p = beta p + D^-1 r cg.h:68,75-82
//
// Scalars never visit the host: they live in two CgState records indexed by iteration parity, so
// the kernel that writes the next iteration's record never races with readers of the current one.
// Dot products use the deterministic last-block reduction of reduce.cuh (la::inner_product /
// squared_norm sum owned entries only: cg.h:53,65,74).
#include "comm.h"
#include "envopt.h"
#include "kernels.h"
#include "layout.h"
#include "peer.cuh"
#include "reduce.cuh"
#include "tma.cuh"
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <type_traits>

namespace ptb
{
namespace
{

constexpr int SPMV_THREADS = 256;
// The persistent loop runs ONE 1024-thread CTA per SM: its grid barrier costs one arrival record
// per CTA (148 instead of 592), and the operator phase keeps the same 32 warps per SM.
#ifdef PTB_HOST_EMU
constexpr int LOOP_THREADS = 256; // the host harness plays a CTA with one std::thread per thread
#else
constexpr int LOOP_THREADS = 1024;
#endif
constexpr int VEC_THREADS = 256;

// Gather of the input vector: read-only (non-coherent) path, except for slices that read ghost
// entries in the fused-halo kernel -- those were written by this very launch, so they bypass L1.
enum class Ld
{
  NC,
  CG,
  CA, // coherent cached load (ld.global.ca): the vector changes inside a persistent kernel, L1
      // lines are dropped by the acquire of the grid barrier, never served from the .nc path
  MIX // ghost-reading slices of the persistent loop: owned entries through L1 (.ca), ghost entries
      // (pulled by other CTAs of this very launch, index >= n_rows * BS) past it (.cg); a plain
      // ldp<MIX> is the .ca load (used for the row's own entry, which is owned)
};
template <Ld L>
__device__ __forceinline__ double ldp(const double* q)
{
  if constexpr (L == Ld::NC)
    return __ldg(q);
  else if constexpr (L == Ld::CG)
    return __ldcg(q);
  else
    return ld_ca_f64(q);
}
// the same with the entry's index at hand (Ld::MIX needs it)
template <Ld L>
__device__ __forceinline__ double ldp_at(const double* base, std::int64_t i, std::int64_t ghost_from)
{
  if constexpr (L == Ld::MIX)
    return i >= ghost_from ? __ldcg(base + i) : ld_ca_f64(base + i);
  else
    return ldp<L>(base + i);
}

// One SELL-32 slice: y[row] = sum_k vals * p[col]; returns this row's p.y contribution.
// Matrix values and column indices go through ld_stream with the slice's L2 policy (pinned prefix:
// evict_last, streamed rest: evict_first; sync_ops.cuh), the gathers of p through ldp<L>.
struct L2Plan
{
  unsigned long long stream, pinned;
};
__device__ __forceinline__ L2Plan l2_plan(const SpmvArgs& A)
{
  return A.l2_mode ? L2Plan{l2_policy(1), l2_policy(2)} : L2Plan{l2_policy(0), l2_policy(0)};
}

template <bool POL, typename T>
__device__ __forceinline__ T ldm(const T* q, unsigned long long pol)
{
  if constexpr (POL)
    return ld_stream(q, pol);
  else
    return __ldg(q);
}

// POL = false: plain read-only loads of the matrix (the scalar kernels of large problems, where the
// hints measured no gain: Poisson 20 M DOFs 0.519 vs 0.525 ms).
template <int BS, Ld L, bool POL = true>
__device__ __forceinline__ double spmv_slice(const SpmvArgs& A, const L2Plan& LP,
                                             const double* __restrict__ p, double* __restrict__ y,
                                             std::int32_t slice, int lane)
{
  const std::int64_t mo = A.mat_off[slice];
  const unsigned long long pol = mo < A.pin_entries ? LP.pinned : LP.stream;
  const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
  const std::int32_t row = slice * 32 + lane;
  if constexpr (BS == 1)
  {
    const double* __restrict__ vp = A.vals + mo + lane;
    // Compressed columns (layout.h): one delta per (slice, k) when the stencil is translation
    // invariant over the slice, explicit indices otherwise. The deltas of up to 32 entries are
    // fetched with ONE coalesced load and broadcast by shuffle, so neither the value loads nor
    // the gathers of p wait on index loads.
    const std::int32_t* __restrict__ dp = A.cdelta + (mo >> 5);
    const std::int32_t* __restrict__ xp = A.colsx + A.xoff[slice] + lane;
    double sum = 0.0;
    for (int k0 = 0; k0 < w; k0 += 32)
    {
      const int kn = min(32, w - k0);
      const std::int32_t dl = lane < kn ? ldm<POL>(dp + k0 + lane, pol) : 0;
      const unsigned int em = __ballot_sync(0xffffffffu, lane < kn && dl == INT32_MIN);
      const double* __restrict__ v = vp + static_cast<std::int64_t>(k0) * 32;
      if (em == 0u)
      {
        int kk = 0;
        constexpr int U = 5; // 15 entries per interior P1 row = 3 batches of 5
        for (; kk + U <= kn; kk += U)
        {
          double vv[U], pp[U];
#pragma unroll
          for (int u = 0; u < U; ++u)
          {
            vv[u] = ldm<POL>(v + (kk + u) * 32, pol);
            pp[u] = ldp<L>(p + (row + __shfl_sync(0xffffffffu, dl, kk + u)));
          }
#pragma unroll
          for (int u = 0; u < U; ++u)
            sum += vv[u] * pp[u];
        }
        for (; kk < kn; ++kk)
          sum += ldm<POL>(v + kk * 32, pol) * ldp<L>(p + (row + __shfl_sync(0xffffffffu, dl, kk)));
      }
      else
      {
        // mixed slice: explicit indices sit at xp[rank * 32], rank = number of explicit entries
        // before kk (from the ballot mask), so the four loads of a batch are independent
        int kk = 0;
        for (; kk + 4 <= kn; kk += 4)
        {
          std::int32_t c[4];
          double vv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
          {
            const std::int32_t d = __shfl_sync(0xffffffffu, dl, kk + u);
            const int rank = __popc(em & ((1u << (kk + u)) - 1u));
            c[u] = ((em >> (kk + u)) & 1u) ? ldm<POL>(xp + rank * 32, pol) : row + d;
            vv[u] = ldm<POL>(v + (kk + u) * 32, pol);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            sum += vv[u] * ldp<L>(p + c[u]);
        }
        for (; kk < kn; ++kk)
        {
          const std::int32_t d = __shfl_sync(0xffffffffu, dl, kk);
          const int rank = __popc(em & ((1u << kk) - 1u));
          const std::int32_t c = ((em >> kk) & 1u) ? ldm<POL>(xp + rank * 32, pol) : row + d;
          sum += ldm<POL>(v + kk * 32, pol) * ldp<L>(p + c);
        }
        xp += __popc(em) * 32;
      }
    }
    if (row < A.n_rows)
    {
      y[row] = sum;
      return sum * ldp<L>(p + row);
    }
    return 0.0;
  }
  else
  {
    const std::int32_t* __restrict__ cp = A.cols + mo + lane;
    const double* __restrict__ vp = A.vals + mo * 9 + lane;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < w; ++k)
    {
      const std::int64_t c = ldm<POL>(cp + k * 32, pol);
      const double* __restrict__ v = vp + static_cast<std::int64_t>(k) * 9 * 32;
      double a[9];
#pragma unroll
      for (int e = 0; e < 9; ++e)
        a[e] = ldm<POL>(v + e * 32, pol);
      const double p0 = ldp<L>(p + 3 * c), p1 = ldp<L>(p + 3 * c + 1), p2 = ldp<L>(p + 3 * c + 2);
      s0 += a[0] * p0 + a[1] * p1 + a[2] * p2;
      s1 += a[3] * p0 + a[4] * p1 + a[5] * p2;
      s2 += a[6] * p0 + a[7] * p1 + a[8] * p2;
    }
    if (row < A.n_rows)
    {
      const std::int64_t r3 = 3 * static_cast<std::int64_t>(row);
      y[r3] = s0, y[r3 + 1] = s1, y[r3 + 2] = s2;
      return s0 * ldp<L>(p + r3) + s1 * ldp<L>(p + r3 + 1) + s2 * ldp<L>(p + r3 + 2);
    }
    return 0.0;
  }
}

// ------------------------------------------------------------------------------------------
// Balanced operator application for SMALL per-GPU problems (strong scaling). With whole slices as
// work items a warp gets 2.2 slices on average at 1.25 M DOFs per GPU: 20 % of the warps carry a
// third slice and the kernel spends its last third with a fifth of its warps
// (profiles/r02/ncu_spmv_elasticity_1250k.csv: 51 % warps active against 62.5 % launched, DRAM at
// 57 % of peak where the 10 M-DOF launch reaches 78 %). Here the work unit is one k-step of a slice
// (32 rows x one stored entry): every CTA owns a contiguous run of slices chosen on the host so
// that all CTAs hold the same number of units (+- one slice), and the CTA's warps cut that run
// into equal unit ranges, splitting slices where the cuts fall. A warp that starts inside a slice
// leaves its partial row sums in shared memory; the warp that holds the slice's first entries
// adds the partials in warp order after a CTA barrier, stores y and takes the p.y contribution.
// Same arithmetic per stored entry, fixed order, no atomics; a split row is summed in (at most
// eight) pieces instead of one, so y may differ from the slice-per-warp kernel in the last bit.
// ------------------------------------------------------------------------------------------
constexpr int BAL_SLICES_PER_WARP = 8; // the shared-memory prefix holds 8 slices per warp of the CTA

// Partial row sums of one slice over its entries [kb, ke).
template <int BS, Ld L>
__device__ __forceinline__ void spmv_slice_part(const SpmvArgs& A, const L2Plan& LP,
                                                const double* __restrict__ p, std::int32_t slice,
                                                int kb, int ke, int lane, double (&s)[BS])
{
  const std::int64_t mo = A.mat_off[slice];
  const unsigned long long pol = mo < A.pin_entries ? LP.pinned : LP.stream;
  const std::int32_t row = slice * 32 + lane;
  const std::int64_t ghost_from = static_cast<std::int64_t>(A.n_rows) * BS; // Ld::MIX
  if constexpr (BS == 1)
  {
    // rows of at most 32 stored entries (P1): one batch of column deltas
    const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
    const double* __restrict__ v = A.vals + mo + lane;
    const std::int32_t dl = lane < w ? ld_stream(A.cdelta + (mo >> 5) + lane, pol) : 0;
    const unsigned int em = __ballot_sync(0xffffffffu, lane < w && dl == INT32_MIN);
    const std::int32_t* __restrict__ xp = A.colsx + A.xoff[slice] + lane;
    double sum = 0.0;
    int kk = kb;
    for (; kk + 4 <= ke; kk += 4)
    {
      std::int32_t c[4];
      double vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
      {
        const std::int32_t d = __shfl_sync(0xffffffffu, dl, kk + u);
        const int rank = __popc(em & ((1u << (kk + u)) - 1u));
        c[u] = ((em >> (kk + u)) & 1u) ? ld_stream(xp + rank * 32, pol) : row + d;
        vv[u] = ld_stream(v + (kk + u) * 32, pol);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        sum += vv[u] * ldp_at<L>(p, c[u], ghost_from);
    }
    for (; kk < ke; ++kk)
    {
      const std::int32_t d = __shfl_sync(0xffffffffu, dl, kk);
      const int rank = __popc(em & ((1u << kk) - 1u));
      const std::int32_t c = ((em >> kk) & 1u) ? ld_stream(xp + rank * 32, pol) : row + d;
      sum += ld_stream(v + kk * 32, pol) * ldp_at<L>(p, c, ghost_from);
    }
    s[0] = sum;
  }
  else
  {
    const std::int32_t* __restrict__ cp = A.cols + mo + lane;
    const double* __restrict__ vp = A.vals + mo * 9 + lane;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int k = kb; k < ke; ++k)
    {
      const std::int64_t c = ld_stream(cp + k * 32, pol);
      const double* __restrict__ v = vp + static_cast<std::int64_t>(k) * 9 * 32;
      double a[9];
#pragma unroll
      for (int e = 0; e < 9; ++e)
        a[e] = ld_stream(v + e * 32, pol);
      const double p0 = ldp_at<L>(p, 3 * c, ghost_from), p1 = ldp_at<L>(p, 3 * c + 1, ghost_from),
                   p2 = ldp_at<L>(p, 3 * c + 2, ghost_from);
      s0 += a[0] * p0 + a[1] * p1 + a[2] * p2;
      s1 += a[3] * p0 + a[4] * p1 + a[5] * p2;
      s2 += a[6] * p0 + a[7] * p1 + a[8] * p2;
    }
    s[0] = s0, s[1] = s1, s[2] = s2;
  }
}

// Shared memory of spmv_cta_balanced, declared once per kernel (the roles call it from two places).
template <int BS, int W>
struct BalShared
{
  std::int32_t su[BAL_SLICES_PER_WARP * W + 1]; // unit prefix of the CTA's slices, from 0
  double head[W][32 * BS];                      // partial sums of the slice a warp starts inside
  std::int32_t head_pos[W];                     // its position in the CTA's run, -1 = none
};

// The slices at positions [i0, i1) of `order`, shared evenly by the warps of this CTA.
// Every thread of the CTA must call this (it contains CTA barriers). Returns the thread's p.y share.
template <int BS, Ld L, int W>
__device__ __forceinline__ double spmv_cta_balanced(const SpmvArgs& A, const L2Plan& LP,
                                                    const double* __restrict__ p,
                                                    double* __restrict__ y,
                                                    const std::int32_t* __restrict__ order,
                                                    std::int32_t i0, std::int32_t i1,
                                                    BalShared<BS, W>& sh, std::int32_t& first_slice,
                                                    int& first_kb)
{
  constexpr int MAXS = BAL_SLICES_PER_WARP * W;
  std::int32_t* su = sh.su;
  double(*head)[32 * BS] = sh.head;
  std::int32_t* head_pos = sh.head_pos;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = i1 - i0;
  __syncthreads(); // the previous use of the shared arrays (persistent loop) is over
  for (int t = threadIdx.x; t <= n; t += blockDim.x)
    su[t] = A.ounit[i0 + t] - A.ounit[i0];
  if (lane == 0)
    head_pos[warp] = -1;
  __syncthreads();
  const std::int32_t total = su[n];
  const std::int32_t u0 = static_cast<std::int32_t>(static_cast<std::int64_t>(total) * warp / W);
  const std::int32_t u1 = static_cast<std::int32_t>(static_cast<std::int64_t>(total) * (warp + 1) / W);
  // position j with su[j] <= u0 < su[j + 1]: the number of t in [1, n] with su[t] <= u0
  int j = 0;
#pragma unroll
  for (int t0 = 0; t0 < MAXS; t0 += 32)
    j += __popc(__ballot_sync(0xffffffffu, t0 + lane < n && su[t0 + lane + 1] <= u0));
  double dotv = 0.0;
  double pend[BS];
  int pend_j = -1;
  std::int32_t u = u0;
  first_slice = -1, first_kb = 0; // where this warp's range starts (the loop's L2 prefetch uses it)
  if (u0 < u1)
    first_slice = order[i0 + j], first_kb = u0 - su[j];
  while (u < u1)
  {
    const int kb = u - su[j];
    const int wj = su[j + 1] - su[j];
    const int ke = min(su[j + 1], u1) - su[j];
    const std::int32_t slice = order[i0 + j];
    double s[BS];
    spmv_slice_part<BS, L>(A, LP, p, slice, kb, ke, lane, s);
    if (kb == 0 && ke == wj)
    {
      const std::int32_t row = slice * 32 + lane;
      if (row < A.n_rows)
      {
#pragma unroll
        for (int a = 0; a < BS; ++a)
        {
          y[static_cast<std::int64_t>(row) * BS + a] = s[a];
          dotv += s[a] * ldp<L>(p + static_cast<std::int64_t>(row) * BS + a);
        }
      }
    }
    else if (kb == 0)
    {
#pragma unroll
      for (int a = 0; a < BS; ++a)
        pend[a] = s[a];
      pend_j = j; // the rest of this slice belongs to the following warps
    }
    else
    {
#pragma unroll
      for (int a = 0; a < BS; ++a)
        head[warp][lane * BS + a] = s[a];
      if (lane == 0)
        head_pos[warp] = j;
    }
    u = su[j] + ke;
    ++j;
  }
  __syncthreads();
  if (pend_j >= 0)
  {
    for (int w2 = warp + 1; w2 < W; ++w2)
      if (head_pos[w2] == pend_j)
      {
#pragma unroll
        for (int a = 0; a < BS; ++a)
          pend[a] += head[w2][lane * BS + a];
      }
    const std::int32_t row = order[i0 + pend_j] * 32 + lane;
    if (row < A.n_rows)
    {
#pragma unroll
      for (int a = 0; a < BS; ++a)
      {
        y[static_cast<std::int64_t>(row) * BS + a] = pend[a];
        dotv += pend[a] * ldp<L>(p + static_cast<std::int64_t>(row) * BS + a);
      }
    }
  }
  return dotv;
}

// Halo exchange fused into the operator (peer mode). The CTAs split into two roles:
//   pullers (blockIdx < npull)  publish "my p is complete", wait for the neighbours, pull the ghost
//                               values of p straight out of the owners' vectors over NVLink, then
//                               process ALL slices that read ghost columns (they are the only ones
//                               that must wait for the pull);
//   the others                  process the slices without ghost columns and never wait.
// npull is chosen on the host in proportion to the ghost-reading share of the work, so neither
// role is the tail of the kernel.
constexpr int MAX_PULL = 256;
struct FusedHalo
{
  PeerHalo H;
  const std::int32_t* order;      // slice visiting order: interior first, ghost-reading last
  std::int32_t n_interior;        // number of leading slices of `order` without ghost columns
  int npull;                      // puller CTAs (<= MAX_PULL, < gridDim.x)
  unsigned long long epoch;       // halo epoch of this launch
  unsigned long long* ready;      // [MAX_PULL] per-puller completion epochs (local memory)
  double* pw;                     // writable alias of p (ghost part)
};


// The pull of one CTA's share of the receive list (generic Scatterer index lists) for halo epoch
// `epoch` (the persistent loop advances the epoch per iteration).
__device__ __forceinline__ void halo_pull_share_at(const PeerView& P, const FusedHalo& FH,
                                                   unsigned long long epoch)
{
  const PeerHalo& H = FH.H;
  if (blockIdx.x == 0 && threadIdx.x < H.n_nbr)
  {
    __threadfence_system(); // p was completed by the previous kernel: publish "ready"
    st_release_sys(&P.win[H.nbr_rank[threadIdx.x]]->halo_flag[P.rank], epoch);
  }
  if (threadIdx.x < H.n_nbr)
  {
    const unsigned long long* flag = &P.win[P.rank]->halo_flag[H.nbr_rank[threadIdx.x]];
    while (ld_acquire_sys(flag) < epoch)
    {
    }
  }
  __syncthreads();
  constexpr int PULL_ILP = 4; // independent remote loads in flight per thread (NVLink ~2 us)
  const std::int64_t n = static_cast<std::int64_t>(H.recv_displ[H.n_nbr]) * H.bs;
  const std::int64_t step = static_cast<std::int64_t>(FH.npull) * blockDim.x;
  for (std::int64_t i0 = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x; i0 < n;
       i0 += step * PULL_ILP)
  {
    double val[PULL_ILP];
    std::int64_t dst[PULL_ILP];
#pragma unroll
    for (int u = 0; u < PULL_ILP; ++u)
    {
      const std::int64_t i = i0 + u * step;
      dst[u] = -1;
      if (i < n)
      {
        const std::int32_t j = static_cast<std::int32_t>(i / H.bs);
        const std::int32_t c = static_cast<std::int32_t>(i - static_cast<std::int64_t>(j) * H.bs);
        int nb = 0;
        while (j >= H.recv_displ[nb + 1])
          ++nb;
        dst[u] = static_cast<std::int64_t>(H.remote_indices[j]) * H.bs + c;
        val[u] = __ldcv(H.peer_p[nb] + static_cast<std::int64_t>(H.src_index[j]) * H.bs + c);
      }
    }
#pragma unroll
    for (int u = 0; u < PULL_ILP; ++u)
      if (dst[u] >= 0)
        FH.pw[dst[u]] = val[u];
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    __threadfence();
    st_release_gpu(&FH.ready[blockIdx.x], epoch);
  }
}

__device__ __forceinline__ void halo_pull_share(const PeerView& P, const FusedHalo& FH)
{
  halo_pull_share_at(P, FH, FH.epoch);
}

template <int BS, bool FUSED, bool BAL = false>
__global__ void __launch_bounds__(SPMV_THREADS, BAL ? 4 : 5)
spmv_sell(SpmvArgs A, const double* __restrict__ p, double* __restrict__ y, CgState* st,
          double* partials, unsigned int* ticket, PeerView P, unsigned int epoch, FusedHalo FH)
{
  __shared__ double red[32];
  if (st != nullptr && st->conv)
    return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int warps_per_cta = SPMV_THREADS / 32;
  const L2Plan LP = l2_plan(A);
  double dotv = 0.0;
  constexpr bool balanced = BAL; // the host passes A.bal_begin / A.ounit with this instantiation
  __shared__ BalShared<BS, BAL ? SPMV_THREADS / 32 : 1> bal_sh;
  [[maybe_unused]] std::int32_t pf_slice = -1;
  [[maybe_unused]] int pf_kb = 0;
  if constexpr (FUSED)
  {
    if (blockIdx.x < FH.npull)
    {
      halo_pull_share(P, FH);
      // all shares must have landed before any ghost column is read
      for (int base = 0; base < FH.npull; base += 32)
      {
        unsigned long long f;
        do
          f = base + lane < FH.npull ? ld_acquire_gpu(&FH.ready[base + lane]) : ~0ull;
        while (!__all_sync(0xffffffffu, f >= FH.epoch));
      }
      if constexpr (balanced)
        dotv = spmv_cta_balanced<BS, Ld::CG, SPMV_THREADS / 32>(A, LP, p, y, FH.order, A.bal_begin[blockIdx.x],
                                             A.bal_begin[blockIdx.x + 1], bal_sh, pf_slice, pf_kb);
      else
        for (std::int32_t it = FH.n_interior + blockIdx.x * warps_per_cta + warp; it < A.n_slices;
             it += FH.npull * warps_per_cta)
          dotv += spmv_slice<BS, Ld::CG, BS == 3>(A, LP, p, y, FH.order[it], lane);
    }
    else if constexpr (balanced)
    {
      const int b = FH.npull + 1 + (blockIdx.x - FH.npull);
      dotv = spmv_cta_balanced<BS, Ld::NC, SPMV_THREADS / 32>(A, LP, p, y, FH.order, A.bal_begin[b], A.bal_begin[b + 1], bal_sh, pf_slice, pf_kb);
    }
    else
    {
      const std::int32_t stride = (gridDim.x - FH.npull) * warps_per_cta;
      for (std::int32_t it = (blockIdx.x - FH.npull) * warps_per_cta + warp; it < FH.n_interior;
           it += stride)
        dotv += spmv_slice<BS, Ld::NC, BS == 3>(A, LP, p, y, FH.order[it], lane);
    }
  }
  else if constexpr (balanced)
    dotv = spmv_cta_balanced<BS, Ld::NC, SPMV_THREADS / 32>(A, LP, p, y, FH.order, A.bal_begin[blockIdx.x],
                                         A.bal_begin[blockIdx.x + 1], bal_sh, pf_slice, pf_kb);
  else
  {
    const std::int32_t stride = gridDim.x * warps_per_cta;
    for (std::int32_t it = blockIdx.x * warps_per_cta + warp; it < A.n_slices; it += stride)
      dotv += spmv_slice<BS, Ld::NC, BS == 3>(A, LP, p, y, FH.order[it], lane);
  }
  if (st != nullptr)
  {
    double v[1] = {dotv}, out[1];
    if (grid_sum_last_block<1>(v, partials, ticket, red, out) && threadIdx.x == 0)
    {
      if (P.nranks > 1)
        peer_publish(P, epoch, out[0], 0.0); // all-reduce of p.y: every rank gets every partial
      else
        st->py = out[0];
    }
  }
}

#ifndef PTB_HOST_EMU // TMA / mbarrier PTX: device build only
// ------------------------------------------------------------------------------------------
// TMA-staged variant of the scalar operator (rows with <= 32 stored entries, i.e. P1).
// With the column indices compressed away the kernel streams almost nothing but matrix values,
// and a thread-per-row loop cannot keep enough of them in flight from registers. Here every warp
// owns a ring of TMA_STAGES shared-memory buffers; lane 0 fetches whole slices (w x 256 B of
// values, one contiguous block in SELL-32) with cp.async.bulk (TMA, SASS UBLKCP) completing on an
// mbarrier, TMA_STAGES slices ahead, while the warp gathers p for the current slice from
// registers. Bytes in flight per SM are set by the ring depth, not by the register file.
// ------------------------------------------------------------------------------------------
constexpr int TMA_STAGES = 3;
constexpr int TMA_THREADS = 256;

__global__ void __launch_bounds__(TMA_THREADS, 2)
spmv_sell_tma(SpmvArgs A, const double* __restrict__ p, double* __restrict__ y, CgState* st,
              double* partials, unsigned int* ticket, PeerView P, unsigned int epoch, FusedHalo FH,
              int stage_doubles)
{
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ double red[32];
  if (st != nullptr && st->conv)
    return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int warps_per_cta = TMA_THREADS / 32;
  const std::int32_t warp0 = blockIdx.x * warps_per_cta + warp;
  const std::int32_t stride = gridDim.x * warps_per_cta;
  double* ring = reinterpret_cast<double*>(dsm) + static_cast<std::size_t>(warp) * TMA_STAGES * stage_doubles;
  std::uint64_t* bars = reinterpret_cast<std::uint64_t*>(
                            dsm + static_cast<std::size_t>(warps_per_cta) * TMA_STAGES * stage_doubles * sizeof(double))
                        + warp * TMA_STAGES;
  if (lane == 0)
  {
#pragma unroll
    for (int s = 0; s < TMA_STAGES; ++s)
      mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  // slices of this warp: it = warp0 + j*stride, j = 0 .. n_my-1
  const std::int32_t n_my = warp0 < A.n_slices ? (A.n_slices - warp0 + stride - 1) / stride : 0;
  auto slice_of = [&](std::int32_t j) -> std::int32_t {
    return FH.order[warp0 + j * stride];
  };
  auto issue = [&](std::int32_t j) {
    if (lane == 0)
    {
      const std::int32_t sl = slice_of(j);
      const std::int64_t mo = A.mat_off[sl];
      const std::uint32_t bytes = static_cast<std::uint32_t>((A.mat_off[sl + 1] - mo) * sizeof(double));
      std::uint64_t* bar = &bars[j % TMA_STAGES];
      mbar_expect_tx(bar, bytes);
      if (bytes > 0)
        tma_load_1d(ring + (j % TMA_STAGES) * stage_doubles, A.vals + mo, bytes, bar);
    }
  };
  for (std::int32_t j = 0; j < min(n_my, TMA_STAGES); ++j)
    issue(j);

  // Software pipeline over this warp's slices: the column deltas are loaded two slices ahead and
  // the gathers of p one slice ahead, so in steady state nothing an iteration consumes was
  // requested in that iteration (values: TMA ring; p: registers; deltas: registers).
  struct Meta
  {
    std::int32_t slice, dl;
    int w;
    bool ghost;
  };
  auto load_meta = [&](std::int32_t j) -> Meta {
    Meta m{0, 0, 0, false};
    if (j < n_my)
    {
      m.slice = slice_of(j);
      const std::int64_t mo = A.mat_off[m.slice];
      m.w = static_cast<int>((A.mat_off[m.slice + 1] - mo) >> 5); // <= 32 by construction
      m.dl = lane < m.w ? __ldg(A.cdelta + (mo >> 5) + lane) : 0;
    }
    return m;
  };
  // fast slices: all columns are row + delta, at most 16 entries, no ghost columns
  auto is_fast = [&](const Meta& m) -> bool {
    const unsigned int em = __ballot_sync(0xffffffffu, lane < m.w && m.dl == INT32_MIN);
    return em == 0u && m.w <= 16 && !m.ghost;
  };
  auto gather = [&](const Meta& m, double (&pp)[16]) {
    const std::int32_t row = m.slice * 32 + lane;
#pragma unroll
    for (int u = 0; u < 16; ++u)
    {
      const std::int32_t d = __shfl_sync(0xffffffffu, m.dl, u);
      pp[u] = u < m.w ? __ldg(p + (row + d)) : 0.0;
    }
  };

  double dotv = 0.0;
  // Ring of 4 slice descriptors and 2 gather buffers, indexed statically (the loop is unrolled by
  // 4): a register that a load is still filling is never moved, so nothing stalls on a copy.
  Meta M[4];
  double PP[2][16];
  bool fast[4] = {false, false, false, false};
  M[0] = load_meta(0);
  M[1] = load_meta(1);
  fast[0] = n_my > 0 && is_fast(M[0]);
  if (fast[0])
    gather(M[0], PP[0]);

  auto step = [&](std::int32_t j, auto S_) {
    constexpr int S = decltype(S_)::value;
    constexpr int S1 = (S + 1) & 3, S2 = (S + 2) & 3;
    if (j >= n_my)
      return;
    // prefetch: gathers of slice j+1, descriptor of slice j+2
    fast[S1] = j + 1 < n_my && is_fast(M[S1]);
    if (fast[S1])
      gather(M[S1], PP[(S + 1) & 1]);
    M[S2] = load_meta(j + 2);

    const Meta& cur = M[S];
    const std::int32_t slice = cur.slice;
    const int w = cur.w;
    const std::int32_t row = slice * 32 + lane;
    const double* __restrict__ v = ring + (j % TMA_STAGES) * stage_doubles + lane;
    std::uint64_t* bar = &bars[j % TMA_STAGES];
    const std::uint32_t parity = (j / TMA_STAGES) & 1;
    double sum = 0.0;
    double prow;
    if (fast[S])
    {
      const double(&pp)[16] = PP[S & 1];
      mbar_wait(bar, parity);
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if (u < w)
          sum += v[u * 32] * pp[u];
      prow = __ldg(p + row);
    }
    else
    {
      const unsigned int em = __ballot_sync(0xffffffffu, lane < w && cur.dl == INT32_MIN);
      const std::int32_t* __restrict__ xp = A.colsx + A.xoff[slice] + lane;
      mbar_wait(bar, parity);
      for (int kk = 0; kk < w; ++kk)
      {
        std::int32_t c = row + __shfl_sync(0xffffffffu, cur.dl, kk);
        if ((em >> kk) & 1u)
        {
          c = xp[0];
          xp += 32;
        }
        sum += v[kk * 32] * (cur.ghost ? __ldcg(p + c) : __ldg(p + c));
      }
      prow = row < A.n_rows ? (cur.ghost ? __ldcg(p + row) : __ldg(p + row)) : 0.0;
    }
    __syncwarp(); // every lane is done with this stage's buffer
    if (j + TMA_STAGES < n_my)
    {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(j + TMA_STAGES);
    }
    if (row < A.n_rows)
    {
      y[row] = sum;
      dotv += sum * prow;
    }
  };
  for (std::int32_t j = 0; j < n_my; j += 4)
  {
    step(j, std::integral_constant<int, 0>{});
    step(j + 1, std::integral_constant<int, 1>{});
    step(j + 2, std::integral_constant<int, 2>{});
    step(j + 3, std::integral_constant<int, 3>{});
  }
  if (st != nullptr)
  {
    double vsum[1] = {dotv}, out[1];
    if (grid_sum_last_block<1>(vsum, partials, ticket, red, out) && threadIdx.x == 0)
    {
      if (P.nranks > 1)
        peer_publish(P, epoch, out[0], 0.0);
      else
        st->py = out[0];
    }
  }
}

#endif // PTB_HOST_EMU

// Global sums for the consumer kernels: one thread per CTA collects the nranks partials from the
// local window (peer mode) or reads the locally reduced / NCCL-reduced values.
__device__ __forceinline__ void global_sums(const PeerView& P, unsigned int epoch, double l0,
                                            double l1, double* sh, double& s0, double& s1)
{
  if (P.nranks > 1)
  {
    if (threadIdx.x == 0)
      peer_collect(P, epoch, sh[0], sh[1]);
    __syncthreads();
    s0 = sh[0], s1 = sh[1];
  }
  else
    s0 = l0, s1 = l1;
}

// r = b - y (cg.h:47), p = z = D^-1 r (cg.h:50), local r.r and r.z.
__global__ void __launch_bounds__(VEC_THREADS)
cg_init(std::int64_t n, const double* __restrict__ b, const double* __restrict__ y,
        const double* __restrict__ dinv, double* __restrict__ r, double* __restrict__ p,
        CgState* st, double* partials, unsigned int* ticket, PeerView P, unsigned int epoch)
{
  __shared__ double red[64];
  double v[2] = {0.0, 0.0};
  for (std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<std::int64_t>(gridDim.x) * blockDim.x)
  {
    const double ri = -1.0 * y[i] + b[i];
    const double zi = dinv[i] * ri;
    r[i] = ri;
    p[i] = zi;
    v[0] += ri * ri;
    v[1] += ri * zi;
  }
  double out[2];
  if (grid_sum_last_block<2>(v, partials, ticket, red, out) && threadIdx.x == 0)
  {
    if (P.nranks > 1)
      peer_publish(P, epoch, out[0], out[1]);
    else
      st->rr = out[0], st->rz = out[1];
  }
}

__global__ void cg_finish_init(CgState* st, double rtol, PeerView P, unsigned int epoch)
{
  // st->rr, st->rz hold the (all-reduced) initial sums (cg.h:53-55)
  if (P.nranks > 1)
    peer_collect(P, epoch, st->rr, st->rz);
  st->rnorm0 = st->rr;
  st->rnorm = st->rr;
  st->rz_old = st->rz;
  st->rtol2 = rtol * rtol;
  st->py = 0.0;
  st->k = 0;
  st->conv = 0;
}

// r -= alpha y (cg.h:71) with the local r.r and r.z (cg.h:74). The x update of cg.h:68 is deferred
// to cg_direction, which streams p anyway: 32 B/DOF here instead of 56.
__global__ void __launch_bounds__(VEC_THREADS)
cg_update(std::int64_t n, const double* __restrict__ y, const double* __restrict__ dinv,
          double* __restrict__ r, CgState* cur, double* partials, unsigned int* ticket, PeerView P,
          unsigned int epoch_in, unsigned int epoch_out)
{
  __shared__ double red[64];
  __shared__ double sh[2];
  if (cur->conv)
    return;
  double py, unused;
  global_sums(P, epoch_in, cur->py, 0.0, sh, py, unused);
  const double alpha = cur->rz_old / py; // cg.h:65
  if (blockIdx.x == 0 && threadIdx.x == 0)
    cur->alpha = alpha; // for cg_direction; no CTA of this kernel reads it
  double v[2] = {0.0, 0.0};
  const std::int64_t n2 = n >> 1;
  const double2* __restrict__ y2 = reinterpret_cast<const double2*>(y);
  const double2* __restrict__ d2 = reinterpret_cast<const double2*>(dinv);
  double2* __restrict__ r2 = reinterpret_cast<double2*>(r);
  for (std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x; i < n2;
       i += static_cast<std::int64_t>(gridDim.x) * blockDim.x)
  {
    const double2 yy = y2[i], dd = d2[i];
    double2 rr = r2[i];
    rr.x = -alpha * yy.x + rr.x;
    rr.y = -alpha * yy.y + rr.y;
    r2[i] = rr;
    v[0] += rr.x * rr.x;
    v[1] += rr.x * (dd.x * rr.x);
    v[0] += rr.y * rr.y;
    v[1] += rr.y * (dd.y * rr.y);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
  {
    const double ri = -alpha * y[n - 1] + r[n - 1];
    r[n - 1] = ri;
    v[0] += ri * ri;
    v[1] += ri * (dinv[n - 1] * ri);
  }
  double out[2];
  if (grid_sum_last_block<2>(v, partials, ticket, red, out) && threadIdx.x == 0)
  {
    if (P.nranks > 1)
      peer_publish(P, epoch_out, out[0], out[1]);
    else
      cur->rr = out[0], cur->rz = out[1];
  }
}

// x += alpha p (cg.h:68, deferred), beta and the stopping rule (cg.h:75-79), p = beta p + D^-1 r
// (cg.h:82) unless converged. 48 B/DOF.
__global__ void __launch_bounds__(VEC_THREADS, 5)
cg_direction(std::int64_t n, const double* __restrict__ r, const double* __restrict__ dinv,
             double* __restrict__ p, double* __restrict__ x, const CgState* cur, CgState* nxt,
             PeerView P, unsigned int epoch)
{
  __shared__ double sh[2];
  const bool first = blockIdx.x == 0 && threadIdx.x == 0;
  if (cur->conv)
  {
    if (first)
      *nxt = *cur;
    return;
  }
  double rr, rz;
  global_sums(P, epoch, cur->rr, cur->rz, sh, rr, rz);
  const double alpha = cur->alpha;
  const double beta = rz / cur->rz_old;                 // cg.h:75
  const bool converged = rr / cur->rnorm0 < cur->rtol2; // cg.h:78
  if (first)
  {
    CgState s = *cur;
    s.rz_old = rz;
    s.rnorm = rr;
    s.k = cur->k + 1;
    s.conv = converged ? 1 : 0;
    *nxt = s;
  }
  const std::int64_t n2 = n >> 1;
  const double2* __restrict__ r2 = reinterpret_cast<const double2*>(r);
  const double2* __restrict__ d2 = reinterpret_cast<const double2*>(dinv);
  double2* __restrict__ p2 = reinterpret_cast<double2*>(p);
  double2* __restrict__ x2 = reinterpret_cast<double2*>(x);
  for (std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x; i < n2;
       i += static_cast<std::int64_t>(gridDim.x) * blockDim.x)
  {
    double2 pp = p2[i], xx = x2[i];
    xx.x = alpha * pp.x + xx.x;
    xx.y = alpha * pp.y + xx.y;
    x2[i] = xx;
    if (!converged)
    {
      const double2 rv = r2[i], dd = d2[i];
      pp.x = beta * pp.x + dd.x * rv.x;
      pp.y = beta * pp.y + dd.y * rv.y;
      p2[i] = pp;
    }
  }
  if ((n & 1) && first)
  {
    const double pv = p[n - 1];
    x[n - 1] = alpha * pv + x[n - 1];
    if (!converged)
      p[n - 1] = beta * pv + dinv[n - 1] * r[n - 1];
  }
}

__global__ void publish_py(const CgState* st, PeerView P, unsigned int epoch)
{
  if (!st->conv)
    peer_publish(P, epoch, st->py, 0.0);
}

__global__ void fill_kernel(double* v, std::int64_t n, double value)
{
  for (std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<std::int64_t>(gridDim.x) * blockDim.x)
    v[i] = value;
}

// pack_fn / unpack_fn of cgpoisson_problem.cpp:32-44 (block indices, bs values each)
__global__ void pack_kernel(const double* __restrict__ v, const std::int32_t* __restrict__ idx,
                            std::int64_t n, int bs, double* __restrict__ out)
{
  const std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (i < n * bs)
    out[i] = v[static_cast<std::int64_t>(idx[i / bs]) * bs + i % bs];
}
__global__ void unpack_kernel(const double* __restrict__ in, const std::int32_t* __restrict__ idx,
                              std::int64_t n, int bs, double* __restrict__ v)
{
  const std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (i < n * bs)
    v[static_cast<std::int64_t>(idx[i / bs]) * bs + i % bs] = in[i];
}

__global__ void __launch_bounds__(VEC_THREADS)
sqnorm_kernel(std::int64_t n, const double* __restrict__ v, double* out, double* partials,
              unsigned int* ticket, PeerView P, unsigned int epoch)
{
  __shared__ double red[32];
  double s[1] = {0.0};
  for (std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<std::int64_t>(gridDim.x) * blockDim.x)
    s[0] += v[i] * v[i];
  double o[1];
  if (grid_sum_last_block<1>(s, partials, ticket, red, o) && threadIdx.x == 0)
  {
    if (P.nranks > 1)
    {
      double t0, t1;
      peer_publish(P, epoch, o[0], 0.0);
      peer_collect(P, epoch, t0, t1);
      o[0] = t0;
    }
    *out = o[0];
  }
}

// ------------------------------------------------------------------------------------------
// The whole iteration loop of cg.h:57-84 in ONE cooperative kernel (opt-in, PTB_CG_PERSISTENT=1).
// Same three phases and the same arithmetic as spmv_sell / cg_update / cg_direction, separated
// by grid barriers instead of kernel boundaries:
//   phase 1  y = A p (+ fused halo pull in peer mode), CTA partials of p.y
//   barrier  the last CTA to arrive sums the partials in index order, publishes the local sum to
//            the peers (LL window) and releases the barrier; every CTA collects the global sum
//   phase 2  alpha, r -= alpha y, CTA partials of r.r and r.z
//   barrier  as above (two values)
//   phase 3  beta, stopping rule (cg.h:78), x += alpha p, p = beta p + D^-1 r
//   barrier  p complete before the next operator application
// Why: at 8 GPUs (elasticity, 1.25 M DOFs per GPU) an iteration takes 155 us against 97 us of
// kernel time; the 7-10 us vector kernels pay a launch + drain + fill each, and every kernel
// boundary adds to the skew the two all-reduces then wait for. Here the host launches once per
// solve and reads the iteration count at the end.
// Vectors that change inside the kernel (p, r, x, y) are never read through the non-coherent
// path: gathers of p use ld.global.ca / .cg, the streaming phases use .cg loads.
// Default for block rows up to 3 M DOFs per GPU since round 2 (profiles/r02/loop_trace_history.txt).
// ------------------------------------------------------------------------------------------
struct LoopArgs
{
  SpmvArgs A;
  std::int64_t n; // owned entries
  const double* dinv;
  double *r, *p, *x, *y;
  CgState* st;               // [2], indexed by iteration parity
  unsigned long long* slots; // [gridDim.x + 1][4] LL records: one arrival record per CTA + the release
  int it0, n_it;             // iterations it0+1 .. it0+n_it
  unsigned int ebase;        // peer reduction epochs: ebase + 2 (j-1) for p.y, + 1 for (r.r, r.z)
  unsigned int lbase;        // grid barrier epochs: lbase + 3 (j-1) + {0, 1, 2}
  // phase trace (PTB_LOOP_TRACE=iteration): [gridDim.x][8] globaltimer stamps of thread 0 of every
  // CTA in iteration trace_iter: start, SpMV done, barrier 1 passed, update done, barrier 2 passed,
  // direction done, barrier 3 passed. nullptr = no trace (one predictable branch per phase).
  unsigned long long* trace;
  int trace_iter;
  int pf_bytes; // bytes of its matrix range every warp prefetches into L2 per iteration (0 = off)
  // Resident vectors (balanced instantiation, runs of at most res_cap / (32 BS) slices; opt-in,
  // measured slower, see launch_cg_loop): x and r of the CTA's own rows live in dynamic shared memory
  // for the whole solve, which takes 40 of the vector phases' 96 B/DOF out of the L2 traffic.
  // res_cap = entries per vector (0 = off); x and r are written back when the loop ends.
  int res_cap;
};

// While the vectors are updated (L2-resident under strong scaling) and the CTAs wait in the grid
// barriers, HBM idles: ~20 us of a ~105 us iteration at 1.25 M DOFs per GPU. Every warp uses that time
// to pull the first pf_bytes of ITS range of the matrix (the same addresses every iteration) into L2
// with prefetch.global.L2, one 128-byte line per lane and step; the next operator phase then starts on
// L2 hits. [byte0, byte0 + nbytes) of the range that starts at entry kb of `slice`.
template <int BS>
__device__ __forceinline__ void prefetch_matrix_run(const SpmvArgs& A, std::int32_t slice, int kb, int lane,
                                                    int byte0, int nbytes)
{
#ifndef PTB_HOST_EMU
  if (slice < 0 || nbytes <= 0)
    return;
  const std::int64_t e0 = A.mat_off[slice] + static_cast<std::int64_t>(kb) * 32, e1 = A.mat_off[A.n_slices];
  const char* v = reinterpret_cast<const char*>(A.vals + e0 * (BS * BS));
  const char* vend = reinterpret_cast<const char*>(A.vals + e1 * (BS * BS));
  for (int off = byte0 + lane * 128; off < byte0 + nbytes; off += 32 * 128)
    if (v + off < vend)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(v + off));
  if constexpr (BS == 3)
  {
    // the column indices of the same entries: 4 bytes per 72 bytes of values
    const char* c = reinterpret_cast<const char*>(A.cols + e0);
    const char* cend = reinterpret_cast<const char*>(A.cols + e1);
    const int cb0 = byte0 / 18, cn = nbytes / 18;
    for (int off = cb0 + lane * 128; off < cb0 + cn; off += 32 * 128)
      if (c + off < cend)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
  }
#endif
}

__device__ __forceinline__ void loop_stamp(const LoopArgs& L, int j, int slot)
{
#ifndef PTB_HOST_EMU
  if (L.trace != nullptr && j == L.trace_iter && threadIdx.x == 0)
  {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    L.trace[static_cast<std::size_t>(blockIdx.x) * 8 + slot] = t;
  }
#endif
}


// Grid barrier fused with a deterministic reduction of NV values per CTA (NV = 0: barrier only).
// No atomics and no leader: every CTA stores one LL arrival record (its partial sums + the barrier
// epoch in the same 8-byte words) and then reads ALL records -- each thread polls its share with
// the loads of all its records in flight at once -- and adds them in index order with the fixed
// block tree, so every CTA of the grid computes the same bits without a second hop. Across GPUs
// CTA 0 publishes the local sums to the peers' windows and every CTA collects the ranks' sums from
// the local window (rank order). On return every thread holds the (peer-)global sums in out[].
// History (profiles/r02/loop_trace_*.txt): 592 same-address atomics + last-CTA reduce + release
// flag ~8 us per barrier; arrival records collected by CTA 0 + release record 5.2 us; this form
// has one L2 round trip after the last arrival.
constexpr int BAR_MAX_RECORDS = 1; // records per polling thread: grids up to blockDim.x CTAs
template <int NV>
__device__ __forceinline__ void grid_reduce_sync(double (&v)[NV > 0 ? NV : 1], const LoopArgs& L,
                                                 const PeerView& P, unsigned int lepoch,
                                                 unsigned int pepoch, double* red,
                                                 double (&out)[NV > 0 ? NV : 1])
{
  __shared__ double bsum[2][2]; // double-buffered by barrier parity: no trailing CTA barrier
  // Thread 0 may only arrive for the CTA once every thread of the CTA has finished the phase:
  // block_sum synchronises the CTA on its way; the plain barrier has to do it itself. (Found by
  // the host harness, tests/emu: without it a neighbour could pull p while it was being written.)
  if constexpr (NV > 0)
    block_sum<NV>(v, red);
  else
    __syncthreads();
  const bool peers = NV > 0 && P.nranks > 1;
  constexpr int NW = NV == 0 ? 1 : NV == 1 ? 2 : 4; // words of the arrival record in use
  if (threadIdx.x == 0)
  {
    fence_acq_rel_gpu(); // the CTA's writes of this phase are ordered before its arrival record
    ll_write_n<NW>(L.slots + 4 * static_cast<std::size_t>(blockIdx.x), lepoch, NV > 0 ? v[0] : 0.0,
                   NV > 1 ? v[NV - 1] : 0.0);
  }
  double a[BAR_MAX_RECORDS], b[BAR_MAX_RECORDS];
  unsigned int todo = 0;
#pragma unroll
  for (int i = 0; i < BAR_MAX_RECORDS; ++i)
    if (threadIdx.x + i * blockDim.x < gridDim.x)
      todo |= 1u << i;
  while (todo != 0u)
  {
#pragma unroll
    for (int i = 0; i < BAR_MAX_RECORDS; ++i)
      if ((todo >> i) & 1u)
      {
        if (ll_try_read_n<NW>(L.slots + 4 * static_cast<std::size_t>(threadIdx.x + i * blockDim.x), lepoch, a[i], b[i]))
          todo &= ~(1u << i);
      }
  }
  double acc[2] = {0.0, 0.0};
  if constexpr (NV > 0)
  {
#pragma unroll
    for (int i = 0; i < BAR_MAX_RECORDS; ++i)
      if (threadIdx.x + i * blockDim.x < gridDim.x)
        acc[0] += a[i], acc[1] += b[i];
  }
  __syncthreads(); // every record has been seen; red is free again
  if constexpr (NV > 0)
    block_sum<2>(acc, red);
  const int par = lepoch & 1u;
  if (threadIdx.x == 0)
  {
    double s0 = acc[0], s1 = acc[1];
    if (peers)
    {
      if (blockIdx.x == 0)
        peer_publish(P, pepoch, s0, s1);
      peer_collect(P, pepoch, s0, s1); // rank order; contains this rank's own slot
    }
    fence_acq_rel_gpu(); // acquire: later loads (and the L1) must not see pre-barrier data
    bsum[par][0] = s0, bsum[par][1] = s1;
  }
  __syncthreads();
  if constexpr (NV > 0)
  {
    out[0] = bsum[par][0];
    if constexpr (NV > 1)
      out[NV - 1] = bsum[par][1];
  }
}

template <int BS, bool FUSED, bool BAL = false>
__global__ void __launch_bounds__(LOOP_THREADS, LOOP_THREADS == 1024 ? 1 : 4)
cg_loop(LoopArgs L, PeerView P, FusedHalo FH)
{
  __shared__ double red[64];
  __shared__ BalShared<BS, BAL ? LOOP_THREADS / 32 : 1> bal_sh;
#ifdef PTB_HOST_EMU
  static double loop_dsm[1 << 16];
#else
  extern __shared__ __align__(16) double loop_dsm[];
#endif
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int warps_per_cta = LOOP_THREADS / 32;
  const SpmvArgs& A = L.A;
  const L2Plan LP = l2_plan(A);
  const bool first = blockIdx.x == 0 && threadIdx.x == 0;
  const std::int64_t n2 = L.n >> 1;
  const std::int64_t tid = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  const std::int64_t nthr = static_cast<std::int64_t>(gridDim.x) * blockDim.x;
  const unsigned long long halo0 = FH.epoch; // epoch of the launch that precedes this loop
  [[maybe_unused]] std::int32_t pf_slice = -1; // start of this warp's matrix range (balanced split)
  [[maybe_unused]] int pf_kb = 0;

  // ---- resident x and r: this CTA's run of slices (the same every iteration) ------------------
  constexpr int ROW = 32 * BS; // entries of one slice
  [[maybe_unused]] double* x_sh = loop_dsm;
  [[maybe_unused]] double* r_sh = loop_dsm + L.res_cap;
  [[maybe_unused]] std::int32_t run0 = 0;
  [[maybe_unused]] int run_entries = 0; // 0: vectors stay in global memory
  if constexpr (BAL)
  {
    if (L.res_cap > 0)
    {
      int b = blockIdx.x;
      if constexpr (FUSED)
        b = blockIdx.x < FH.npull ? blockIdx.x : FH.npull + 1 + (blockIdx.x - FH.npull);
      run0 = A.bal_begin[b];
      run_entries = (A.bal_begin[b + 1] - run0) * ROW;
      for (int idx = threadIdx.x; idx < run_entries; idx += blockDim.x)
      {
        const std::int64_t g = static_cast<std::int64_t>(FH.order[run0 + idx / ROW]) * ROW + idx % ROW;
        x_sh[idx] = g < L.n ? __ldcg(L.x + g) : 0.0;
        r_sh[idx] = g < L.n ? __ldcg(L.r + g) : 0.0;
      }
      __syncthreads();
    }
  }

  for (int j = 1; j <= L.n_it; ++j)
  {
    const int it = L.it0 + j;
    const CgState* cur = &L.st[it & 1];
    CgState* nxt = &L.st[(it + 1) & 1];
    // cur was completed before the last barrier of the previous iteration (or by the host-side
    // init kernels): every CTA reads the same values and takes the same exit
    if (__ldcg(&cur->conv) != 0)
      break;
    const double rz_old = __ldcg(&cur->rz_old), rnorm0 = __ldcg(&cur->rnorm0),
                 rtol2 = __ldcg(&cur->rtol2);
    const int k_done = __ldcg(&cur->k);

    // ---- phase 1: y = A p, local p.y --------------------------------------------------------
    loop_stamp(L, j, 0);
    double dotv = 0.0;
    constexpr bool balanced = BAL;
    if constexpr (FUSED)
    {
      const unsigned long long hep = halo0 + static_cast<unsigned long long>(j);
      if (blockIdx.x < FH.npull)
      {
        halo_pull_share_at(P, FH, hep);
        for (int base = 0; base < FH.npull; base += 32)
        {
          unsigned long long f;
          do
            f = base + lane < FH.npull ? ld_acquire_gpu(&FH.ready[base + lane]) : ~0ull;
          while (!__all_sync(0xffffffffu, f >= hep));
        }
        if constexpr (balanced)
          dotv = spmv_cta_balanced<BS, Ld::MIX, LOOP_THREADS / 32>(A, LP, L.p, L.y, FH.order, A.bal_begin[blockIdx.x],
                                                                A.bal_begin[blockIdx.x + 1], bal_sh, pf_slice, pf_kb);
        else
          for (std::int32_t s = FH.n_interior + blockIdx.x * warps_per_cta + warp; s < A.n_slices;
               s += FH.npull * warps_per_cta)
            dotv += spmv_slice<BS, Ld::CG>(A, LP, L.p, L.y, FH.order[s], lane);
      }
      else if constexpr (balanced)
      {
        const int b = FH.npull + 1 + (blockIdx.x - FH.npull);
        dotv = spmv_cta_balanced<BS, Ld::CA, LOOP_THREADS / 32>(A, LP, L.p, L.y, FH.order, A.bal_begin[b],
                                                             A.bal_begin[b + 1], bal_sh, pf_slice, pf_kb);
      }
      else
      {
        const std::int32_t stride = (gridDim.x - FH.npull) * warps_per_cta;
        for (std::int32_t s = (blockIdx.x - FH.npull) * warps_per_cta + warp; s < FH.n_interior;
             s += stride)
          dotv += spmv_slice<BS, Ld::CA>(A, LP, L.p, L.y, FH.order[s], lane);
      }
    }
    else if constexpr (balanced)
      dotv = spmv_cta_balanced<BS, Ld::CA, LOOP_THREADS / 32>(A, LP, L.p, L.y, FH.order, A.bal_begin[blockIdx.x],
                                           A.bal_begin[blockIdx.x + 1], bal_sh, pf_slice, pf_kb);
    else
    {
      const std::int32_t stride = gridDim.x * warps_per_cta;
      for (std::int32_t s = blockIdx.x * warps_per_cta + warp; s < A.n_slices; s += stride)
        dotv += spmv_slice<BS, Ld::CA>(A, LP, L.p, L.y, FH.order[s], lane);
    }
    const unsigned int ea = L.ebase + 2u * static_cast<unsigned int>(j - 1), eb = ea + 1u;
    const unsigned int la = L.lbase + 3u * static_cast<unsigned int>(j - 1);
    double v1[1] = {dotv}, py[1];
    loop_stamp(L, j, 1);
    grid_reduce_sync<1>(v1, L, P, la, ea, red, py);
    loop_stamp(L, j, 2);
    const double alpha = rz_old / py[0]; // cg.h:65

    // ---- phase 2: r -= alpha y (cg.h:71), local r.r and r.z (cg.h:74) -----------------------
    if constexpr (BAL)
      prefetch_matrix_run<BS>(A, pf_slice, pf_kb, lane, 0, L.pf_bytes / 2);
    double v2[2] = {0.0, 0.0};
    if (BAL && run_entries > 0)
    {
      for (int idx = threadIdx.x; idx < run_entries; idx += blockDim.x)
      {
        const std::int64_t g = static_cast<std::int64_t>(FH.order[run0 + idx / ROW]) * ROW + idx % ROW;
        if (g < L.n)
        {
          const double rr = -alpha * __ldcg(L.y + g) + r_sh[idx];
          r_sh[idx] = rr;
          v2[0] += rr * rr;
          v2[1] += rr * (__ldg(L.dinv + g) * rr);
        }
      }
    }
    else
    {
      const double2* y2 = reinterpret_cast<const double2*>(L.y);
      const double2* d2 = reinterpret_cast<const double2*>(L.dinv);
      double2* r2 = reinterpret_cast<double2*>(L.r);
      for (std::int64_t i = tid; i < n2; i += nthr)
      {
        const double2 yy = __ldcg(y2 + i), dd = __ldg(d2 + i);
        double2 rr = __ldcg(r2 + i);
        rr.x = -alpha * yy.x + rr.x;
        rr.y = -alpha * yy.y + rr.y;
        r2[i] = rr;
        v2[0] += rr.x * rr.x;
        v2[1] += rr.x * (dd.x * rr.x);
        v2[0] += rr.y * rr.y;
        v2[1] += rr.y * (dd.y * rr.y);
      }
      if ((L.n & 1) && first)
      {
        const double ri = -alpha * __ldcg(L.y + L.n - 1) + __ldcg(L.r + L.n - 1);
        L.r[L.n - 1] = ri;
        v2[0] += ri * ri;
        v2[1] += ri * (__ldg(L.dinv + L.n - 1) * ri);
      }
    }
    double rs[2];
    loop_stamp(L, j, 3);
    grid_reduce_sync<2>(v2, L, P, la + 1u, eb, red, rs);
    loop_stamp(L, j, 4);
    const double rr = rs[0], rz = rs[1];
    const double beta = rz / rz_old;             // cg.h:75
    const bool converged = rr / rnorm0 < rtol2;  // cg.h:78
    if (first)
    {
      CgState s;
      s.py = py[0], s.rr = rr, s.rz = rz, s.rz_old = rz, s.rnorm0 = rnorm0, s.rtol2 = rtol2;
      s.rnorm = rr, s.alpha = alpha, s.k = k_done + 1, s.conv = converged ? 1 : 0;
      *nxt = s;
    }

    // ---- phase 3: x += alpha p (cg.h:68), p = beta p + D^-1 r (cg.h:82) ------------------------
    if constexpr (BAL)
      prefetch_matrix_run<BS>(A, pf_slice, pf_kb, lane, L.pf_bytes / 2, L.pf_bytes - L.pf_bytes / 2);
    if (BAL && run_entries > 0)
    {
      for (int idx = threadIdx.x; idx < run_entries; idx += blockDim.x)
      {
        const std::int64_t g = static_cast<std::int64_t>(FH.order[run0 + idx / ROW]) * ROW + idx % ROW;
        if (g < L.n)
        {
          const double pp = __ldcg(L.p + g);
          x_sh[idx] = alpha * pp + x_sh[idx];
          if (!converged)
            L.p[g] = beta * pp + __ldg(L.dinv + g) * r_sh[idx];
        }
      }
    }
    else
    {
      const double2* r2 = reinterpret_cast<const double2*>(L.r);
      const double2* d2 = reinterpret_cast<const double2*>(L.dinv);
      double2* p2 = reinterpret_cast<double2*>(L.p);
      double2* x2 = reinterpret_cast<double2*>(L.x);
      for (std::int64_t i = tid; i < n2; i += nthr)
      {
        double2 pp = __ldcg(p2 + i), xx = __ldcg(x2 + i);
        xx.x = alpha * pp.x + xx.x;
        xx.y = alpha * pp.y + xx.y;
        x2[i] = xx;
        if (!converged)
        {
          const double2 rv = __ldcg(r2 + i), dd = __ldg(d2 + i);
          pp.x = beta * pp.x + dd.x * rv.x;
          pp.y = beta * pp.y + dd.y * rv.y;
          p2[i] = pp;
        }
      }
      if ((L.n & 1) && first)
      {
        const double pv = __ldcg(L.p + L.n - 1);
        L.x[L.n - 1] = alpha * pv + __ldcg(L.x + L.n - 1);
        if (!converged)
          L.p[L.n - 1] = beta * pv + __ldg(L.dinv + L.n - 1) * __ldcg(L.r + L.n - 1);
      }
    }
    double none[1] = {0.0}, none_out[1];
    loop_stamp(L, j, 5);
    grid_reduce_sync<0>(none, L, P, la + 2u, 0u, red, none_out);
    loop_stamp(L, j, 6);
  }
  if (BAL && run_entries > 0)
  {
    // the loop is over (converged or kmax): the solution and the residual go back to global memory
    __syncthreads();
    for (int idx = threadIdx.x; idx < run_entries; idx += blockDim.x)
    {
      const std::int64_t g = static_cast<std::int64_t>(FH.order[run0 + idx / ROW]) * ROW + idx % ROW;
      if (g < L.n)
        L.x[g] = x_sh[idx], L.r[g] = r_sh[idx];
    }
  }
}

#ifndef PTB_HOST_EMU // host launchers: device build only
// Grids are sized to exactly one resident wave: SMs x (CTAs of this kernel that fit on an SM),
// capped by the work. A second partial wave would run at low occupancy and stretch the tail.
template <typename K>
int resident_grid(const ptb_ctx* c, K kernel, int threads, std::int64_t need)
{
  int per_sm = 0;
  PTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
  const std::int64_t cap = static_cast<std::int64_t>(c->num_sms) * std::max(1, std::min(per_sm, 8));
  return static_cast<int>(std::max<std::int64_t>(1, std::min(need, cap)));
}

template <typename K>
int vec_grid(ptb_ctx* c, K kernel, std::int64_t n, int slot)
{
  const std::int64_t need = (n / 2 + VEC_THREADS - 1) / VEC_THREADS;
  if (c->grid_cache[slot] == 0)
    c->grid_cache[slot] = resident_grid(c, kernel, VEC_THREADS, INT32_MAX);
  return static_cast<int>(std::max<std::int64_t>(1, std::min<std::int64_t>(need, c->grid_cache[slot])));
}

} // namespace

int cg_grid(const ptb_ctx* c) { return c->num_sms * 8; }

// One-wave grid of a kernel, cached per context (the occupancy query is a host API call).
template <typename K>
int cached_grid(ptb_ctx* c, int slot, K kernel, int threads, std::size_t smem, std::int64_t need)
{
  if (c->grid_cache[slot] == 0)
  {
    int per_sm = 0;
    PTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    c->grid_cache[slot] = c->num_sms * std::max(1, std::min(per_sm, 8));
    const int forced = env_int("PTB_SPMV_CTAS", 0); // A/B: CTA count of the operator kernels
    if (forced > 0 && (slot <= 4 || (slot >= 16 && slot <= 19)))
      c->grid_cache[slot] = forced;
  }
  return static_cast<int>(std::max<std::int64_t>(1, std::min<std::int64_t>(need, c->grid_cache[slot])));
}

// Host side of the balanced split (spmv_cta_balanced): unit prefix over the slice order and the
// CTAs' runs of positions, equal in units up to one slice. which = 0 spmv_sell, 1 cg_loop; npull < 0:
// no roles. Sets A.ounit / A.bal_begin when the split applies: a problem small enough for slice
// quantisation to matter (fewer than 8 slices per warp), big enough to give every warp work, block
// rows or P1 rows (one batch of column deltas), at most BAL_MAX_SLICES slices per CTA.
void ensure_balance(ptb_ctx* c, SpmvArgs& A, int which, int grid, int npull, int warps_per_cta)
{
  static const bool enabled = env_int("PTB_SPMV_BALANCE", 1) != 0;
  ptb_ctx::Balance& B = c->balance[which];
  const bool compact = c->have_compact;
  if (B.grid != grid || B.npull != npull || B.compact != compact)
  {
    B.grid = grid, B.npull = npull, B.compact = compact, B.ok = false;
    const std::int32_t S = A.n_slices;
    const std::int64_t warps = static_cast<std::int64_t>(grid) * warps_per_cta;
    const int max_run = BAL_SLICES_PER_WARP * warps_per_cta;
    // block rows only: the scalar variant (spmv_slice_part<1>, kept for the tests) lost against the
    // slice-per-warp kernel, whose delta-compressed column path it cannot use across a split
    // (Poisson 500 k DOFs: 30.8 vs 18.3 us, profiles/r02/ab_call7_summary.txt); PTB_SPMV_BALANCE=2 forces it
    static const bool scalar_too = env_int("PTB_SPMV_BALANCE", 1) == 2;
    const bool shape_ok = c->bs == 3 || (scalar_too && c->bs == 1 && c->max_w <= 32 && A.cdelta != nullptr);
    if (enabled && shape_ok && S >= warps && S < 8 * warps && grid >= 1 && (npull < 0 || npull < grid))
    {
      std::vector<std::int64_t> mo(static_cast<std::size_t>(S) + 1);
      std::vector<std::int32_t> order(S);
      PTB_CUDA(cudaStreamSynchronize(c->stream));
      PTB_CUDA(cudaMemcpy(mo.data(), A.mat_off, mo.size() * sizeof(std::int64_t), cudaMemcpyDeviceToHost));
      PTB_CUDA(cudaMemcpy(order.data(), c->slice_order.p, order.size() * sizeof(std::int32_t),
                          cudaMemcpyDeviceToHost));
      std::vector<std::int32_t> ou, begin;
      const int longest = build_balance_plan(mo.data(), order.data(), S, c->n_interior_slices, grid, npull, ou, begin);
      const bool ok = longest <= max_run;
      if (ok)
      {
        B.ounit.upload(ou, c->stream);
        B.begin.upload(begin, c->stream);
        PTB_CUDA(cudaStreamSynchronize(c->stream)); // the host vectors go out of scope
        B.ok = true;
        B.longest_run = longest;
      }
    }
  }
  if (B.ok)
    A.ounit = B.ounit.p, A.bal_begin = B.begin.p;
}

void launch_spmv(ptb_ctx* c, const double* p, double* y, CgState* st, unsigned int epoch,
                 bool fused_halo)
{
  SpmvArgs A = spmv_args(c);
  const std::int64_t need = (c->n_slices + SPMV_THREADS / 32 - 1) / (SPMV_THREADS / 32);
  const PeerView P = peer_view(c);
  FusedHalo FH{};
  FH.order = c->slice_order.p;
  FH.n_interior = c->n_interior_slices;
  static const bool use_tma = [] {
    const char* e = std::getenv("PTB_SPMV_TMA");
    return e && e[0] == '1';
  }();
  // pullers get the ghost-reading slices: size their number to that share of the work (+25 %
  // for the pull itself), at least 8 when the grid allows so the remote loads have parallelism
  auto pullers = [&](int grid) {
    const double share = c->n_slices > 0
                             ? static_cast<double>(c->n_slices - c->n_interior_slices) / c->n_slices
                             : 0.0;
    const int npull = std::max(8, static_cast<int>(std::ceil(1.25 * share * grid)) + 4);
    return std::max(1, std::min(std::min(npull, MAX_PULL), grid / 2));
  };
  int fused_grid = 0;
  if (fused_halo)
  {
    if (!c->peer.enabled || p != c->p.p)
      throw std::runtime_error("fused halo: peer mode and the search direction vector only");
    fused_grid = c->bs == 1 ? cached_grid(c, 0, spmv_sell<1, true>, SPMV_THREADS, 0, need)
                            : cached_grid(c, 1, spmv_sell<3, true>, SPMV_THREADS, 0, need);
    if (fused_grid < 4)
    {
      // too little work to split into roles: separate pull kernel, then the plain operator
      halo_forward(c, c->p.p);
      fused_halo = false;
    }
  }
  if (fused_halo)
  {
    FH.H = peer_halo(c);
    FH.epoch = ++c->peer.halo_epoch;
    FH.ready = c->peer.ready.p;
    FH.pw = c->p.p;
    // small problem: the balanced instantiation (one resident wave of ITS occupancy), see spmv_cta_balanced
    const int gbal = c->bs == 1 ? cached_grid(c, 16, spmv_sell<1, true, true>, SPMV_THREADS, 0, need)
                                : cached_grid(c, 17, spmv_sell<3, true, true>, SPMV_THREADS, 0, need);
    FH.npull = pullers(gbal);
    ensure_balance(c, A, 0, gbal, FH.npull, SPMV_THREADS / 32);
    if (A.bal_begin != nullptr)
    {
      if (c->bs == 1)
        spmv_sell<1, true, true><<<gbal, SPMV_THREADS, 0, c->stream>>>(A, p, y, st, c->partials.p,
                                                                       c->tickets.p, P, epoch, FH);
      else
        spmv_sell<3, true, true><<<gbal, SPMV_THREADS, 0, c->stream>>>(A, p, y, st, c->partials.p,
                                                                       c->tickets.p, P, epoch, FH);
    }
    else
    {
      const int grid = fused_grid;
      FH.npull = pullers(grid);
      if (c->bs == 1)
        spmv_sell<1, true><<<grid, SPMV_THREADS, 0, c->stream>>>(A, p, y, st, c->partials.p,
                                                                 c->tickets.p, P, epoch, FH);
      else
        spmv_sell<3, true><<<grid, SPMV_THREADS, 0, c->stream>>>(A, p, y, st, c->partials.p,
                                                                 c->tickets.p, P, epoch, FH);
    }
  }
  else if (c->bs == 1 && c->max_w <= 32 && use_tma)
  {
    // opt-in TMA-staged scalar kernel (profiles/r01_spmv_ab_*.txt: measured slower than registers)
    const int stage_doubles = std::max(1, c->max_w) * 32;
    const std::size_t smem = static_cast<std::size_t>(TMA_THREADS / 32) * TMA_STAGES
                             * (stage_doubles * sizeof(double) + sizeof(std::uint64_t));
    PTB_CUDA(cudaFuncSetAttribute(spmv_sell_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    const int grid = cached_grid(c, 2, spmv_sell_tma, TMA_THREADS, smem, need);
    spmv_sell_tma<<<grid, TMA_THREADS, smem, c->stream>>>(A, p, y, st, c->partials.p, c->tickets.p,
                                                          P, epoch, FH, stage_doubles);
  }
  else if (c->bs == 1)
  {
    const int gbal = cached_grid(c, 18, spmv_sell<1, false, true>, SPMV_THREADS, 0, need);
    ensure_balance(c, A, 0, gbal, -1, SPMV_THREADS / 32);
    if (A.bal_begin != nullptr)
      spmv_sell<1, false, true><<<gbal, SPMV_THREADS, 0, c->stream>>>(A, p, y, st, c->partials.p,
                                                                      c->tickets.p, P, epoch, FH);
    else
      spmv_sell<1, false><<<cached_grid(c, 3, spmv_sell<1, false>, SPMV_THREADS, 0, need), SPMV_THREADS, 0,
                            c->stream>>>(A, p, y, st, c->partials.p, c->tickets.p, P, epoch, FH);
  }
  else
  {
    const int gbal = cached_grid(c, 19, spmv_sell<3, false, true>, SPMV_THREADS, 0, need);
    ensure_balance(c, A, 0, gbal, -1, SPMV_THREADS / 32);
    if (A.bal_begin != nullptr)
      spmv_sell<3, false, true><<<gbal, SPMV_THREADS, 0, c->stream>>>(A, p, y, st, c->partials.p,
                                                                      c->tickets.p, P, epoch, FH);
    else
      spmv_sell<3, false><<<cached_grid(c, 4, spmv_sell<3, false>, SPMV_THREADS, 0, need), SPMV_THREADS, 0,
                            c->stream>>>(A, p, y, st, c->partials.p, c->tickets.p, P, epoch, FH);
  }
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

bool launch_cg_loop(ptb_ctx* c, const double* dinv, int it0, int n_it, unsigned int ebase,
                    unsigned int lbase, bool fused_halo)
{
  if (c->bs != 1 && c->bs != 3)
    return false;
  LoopArgs L{};
  L.A = spmv_args(c);
  L.n = static_cast<std::int64_t>(c->n_owned) * c->bs;
  L.dinv = dinv;
  L.r = c->r.p, L.p = c->p.p, L.x = c->x.p, L.y = c->y.p;
  L.st = c->cg.p;
  L.it0 = it0, L.n_it = n_it, L.ebase = ebase, L.lbase = lbase;
  PeerView P = peer_view(c);
  FusedHalo FH{};
  FH.order = c->slice_order.p;
  FH.n_interior = c->n_interior_slices;
  const void* kernel = nullptr;
  const void* kernel_bal = nullptr; // the balanced instantiation (spmv_cta_balanced) for small problems
  int slot = 0;
  if (fused_halo)
    kernel = c->bs == 1 ? reinterpret_cast<const void*>(cg_loop<1, true>)
                        : reinterpret_cast<const void*>(cg_loop<3, true>),
    kernel_bal = c->bs == 1 ? reinterpret_cast<const void*>(cg_loop<1, true, true>)
                            : reinterpret_cast<const void*>(cg_loop<3, true, true>),
    slot = c->bs == 1 ? 10 : 11;
  else
    kernel = c->bs == 1 ? reinterpret_cast<const void*>(cg_loop<1, false>)
                        : reinterpret_cast<const void*>(cg_loop<3, false>),
    kernel_bal = c->bs == 1 ? reinterpret_cast<const void*>(cg_loop<1, false, true>)
                            : reinterpret_cast<const void*>(cg_loop<3, false, true>),
    slot = c->bs == 1 ? 12 : 13;
  if (c->grid_cache[slot] == 0)
  {
    // one grid for both instantiations: the smaller of their co-resident capacities
    int per_sm = 0, per_sm_bal = 0;
    PTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, LOOP_THREADS, 0));
    PTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_bal, kernel_bal, LOOP_THREADS, 0));
    c->grid_cache[slot] = c->num_sms * std::max(1, std::min(std::min(per_sm, per_sm_bal), 8));
    const int forced = env_int("PTB_LOOP_CTAS", 0); // A/B: must stay co-resident (cooperative launch)
    if (forced > 0)
      c->grid_cache[slot] = std::min(forced, c->grid_cache[slot]);
  }
  const std::int64_t need = (c->n_slices + LOOP_THREADS / 32 - 1) / (LOOP_THREADS / 32);
  const int grid = static_cast<int>(std::max<std::int64_t>(1, std::min<std::int64_t>(need, c->grid_cache[slot])));
  if (fused_halo)
  {
    if (grid < 4)
      return false; // too little work to split into roles: the caller uses the kernel-per-phase path
    FH.H = peer_halo(c);
    FH.epoch = c->peer.halo_epoch; // iteration j of the loop uses epoch + j
    FH.ready = c->peer.ready.p;
    FH.pw = c->p.p;
    const double share = c->n_slices > 0
                             ? static_cast<double>(c->n_slices - c->n_interior_slices) / c->n_slices
                             : 0.0;
    // 1024-thread CTAs: the ghost-reading share of the slices times PTB_PULL_FACTOR / 100 (default
    // 1.40), at least one. The pullers start ~8 us late (neighbour flag + NVLink round trip) and
    // gather p past L1, ~1.3 us per slice against 1.0 us for an interior slice: with 1.15 they ended
    // the operator phase 15 us after the workers at 8 GPUs (profiles/r02/multi_gpu/trace_8gpu_call12.txt).
    static const double pull_factor = env_int("PTB_PULL_FACTOR", 140) / 100.0;
    const int npull = std::max(LOOP_THREADS >= 1024 ? 1 : 8,
                               static_cast<int>(std::ceil((LOOP_THREADS >= 1024 ? pull_factor : 1.25) * share * grid))
                                   + (LOOP_THREADS >= 1024 ? 0 : 4));
    FH.npull = std::max(1, std::min(std::min(npull, MAX_PULL), grid / 2));
  }
  ensure_balance(c, L.A, 1, grid, fused_halo ? FH.npull : -1, LOOP_THREADS / 32);
  if (grid > BAR_MAX_RECORDS * LOOP_THREADS)
    return false;
  if (c->loop_slots.n < static_cast<std::size_t>(grid + 1) * 4)
  {
    c->loop_slots.alloc(static_cast<std::size_t>(grid + 1) * 4);
    c->loop_slots.zero(c->stream); // epoch 0 = never written; the host counter starts at 1
  }
  L.slots = c->loop_slots.p;
  // L2 prefetch of the matrix during the vector phases (prefetch_matrix_run): KB per warp. Measured
  // and rejected (profiles/r02/ab_call11_summary.txt, elasticity 1.25 M DOFs, us per iteration:
  // 0 KB 109.9, 6 KB 111.0, 12 KB 114.9, 20 KB 117.5, 32 KB 121.7): the operator phase does not get
  // shorter and the vector phases, which run at the L2's bandwidth, get longer. Off by default.
  static const int pf_kb_env = env_int("PTB_LOOP_PREFETCH_KB", 0);
  L.pf_bytes = L.A.bal_begin != nullptr ? std::max(0, pf_kb_env) * 1024 : 0;
  // x and r resident in shared memory when every CTA's run fits (LoopArgs::res_cap). Measured and
  // rejected (profiles/r02/ab_call18_resident.txt, elasticity 1.25 M DOFs): 109.4 us per iteration
  // without, 126.5 us with -- the 190 KB of dynamic shared memory leave ~30 KB of the SM's 256 KB for
  // L1, and the operator phase, whose gathers of p live on L1 hits, grows from 84 to 101 us; the
  // vector phases gain 1 us. Off by default (PTB_LOOP_RESIDENT=1 switches it on).
  static const bool resident_env = env_flag("PTB_LOOP_RESIDENT", false);
  std::size_t dyn_smem = 0;
  L.res_cap = 0;
  if (L.A.bal_begin != nullptr && resident_env)
  {
    const std::size_t cap = static_cast<std::size_t>(c->balance[1].longest_run) * 32 * c->bs;
    cudaFuncAttributes fa{};
    PTB_CUDA(cudaFuncGetAttributes(&fa, kernel_bal));
    if (cap > 0 && 2 * cap * sizeof(double) + fa.sharedSizeBytes + 1024 <= 227u * 1024u)
    {
      dyn_smem = 2 * cap * sizeof(double);
      L.res_cap = static_cast<int>(cap);
      PTB_CUDA(cudaFuncSetAttribute(kernel_bal, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(dyn_smem)));
    }
  }
  static const int trace_iter = env_int("PTB_LOOP_TRACE", 0);
  if (trace_iter > 0)
  {
    c->loop_trace.alloc(static_cast<std::size_t>(grid) * 8);
    c->loop_trace.zero(c->stream);
    L.trace = c->loop_trace.p, L.trace_iter = trace_iter;
  }
  void* args[] = {&L, &P, &FH};
  PTB_CUDA(cudaLaunchCooperativeKernel(L.A.bal_begin != nullptr ? kernel_bal : kernel, dim3(grid),
                                       dim3(LOOP_THREADS), args, dyn_smem, c->stream));
  c->launches += 1;
  if (trace_iter > 0)
  {
    // diagnostic: phase durations of iteration trace_iter over all CTAs (stderr)
    std::vector<unsigned long long> t(static_cast<std::size_t>(grid) * 8);
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    PTB_CUDA(cudaMemcpy(t.data(), c->loop_trace.p, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;
    for (int b = 0; b < grid; ++b)
      if (t[b * 8] != 0)
        t0 = std::min(t0, t[b * 8]);
    const char* names[7] = {"start", "spmv done", "barrier1 out", "update done", "barrier2 out",
                            "direction done", "barrier3 out"};
    std::fprintf(stderr, "[ptb loop trace] iteration %d, %d CTAs, ns after the first CTA's start: min / mean / max\n",
                 trace_iter, grid);
    for (int s2 = 0; s2 < 7; ++s2)
    {
      double mn = 1e30, mx = 0, sum = 0;
      int cnt = 0;
      for (int b = 0; b < grid; ++b)
      {
        if (t[b * 8 + s2] == 0)
          continue;
        const double d = static_cast<double>(t[b * 8 + s2] - t0);
        mn = std::min(mn, d), mx = std::max(mx, d), sum += d, ++cnt;
      }
      if (cnt)
        std::fprintf(stderr, "[ptb loop trace]   %-15s %9.0f %9.0f %9.0f\n", names[s2], mn, sum / cnt, mx);
    }
  }
  return true;
}

void launch_cg_init(ptb_ctx* c, const double* dinv, CgState* st, unsigned int epoch)
{
  const std::int64_t n = static_cast<std::int64_t>(c->n_owned) * c->bs;
  cg_init<<<vec_grid(c, cg_init, 2 * n, 5), VEC_THREADS, 0, c->stream>>>(n, c->b.p, c->y.p, dinv, c->r.p, c->p.p,
                                                         st, c->partials.p, c->tickets.p + 1,
                                                         peer_view(c), epoch);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_cg_finish_init(ptb_ctx* c, CgState* st, double rtol, unsigned int epoch)
{
  cg_finish_init<<<1, 1, 0, c->stream>>>(st, rtol, peer_view(c), epoch);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_cg_update(ptb_ctx* c, const double* dinv, CgState* cur, unsigned int epoch_in,
                      unsigned int epoch_out)
{
  const std::int64_t n = static_cast<std::int64_t>(c->n_owned) * c->bs;
  cg_update<<<vec_grid(c, cg_update, n, 6), VEC_THREADS, 0, c->stream>>>(n, c->y.p, dinv, c->r.p, cur,
                                                           c->partials.p, c->tickets.p + 1,
                                                           peer_view(c), epoch_in, epoch_out);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_cg_direction(ptb_ctx* c, const double* dinv, const CgState* cur, CgState* nxt,
                         unsigned int epoch)
{
  const std::int64_t n = static_cast<std::int64_t>(c->n_owned) * c->bs;
  cg_direction<<<vec_grid(c, cg_direction, n, 7), VEC_THREADS, 0, c->stream>>>(n, c->r.p, dinv, c->p.p, c->x.p, cur,
                                                              nxt, peer_view(c), epoch);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_publish_py(ptb_ctx* c, CgState* st, unsigned int epoch)
{
  if (!c->peer.enabled)
    return;
  publish_py<<<1, 1, 0, c->stream>>>(st, peer_view(c), epoch);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_fill(ptb_ctx* c, double* v, std::int64_t n, double value)
{
  if (n == 0)
    return;
  fill_kernel<<<vec_grid(c, fill_kernel, 2 * n, 8), VEC_THREADS, 0, c->stream>>>(v, n, value);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_pack(ptb_ctx* c, const double* v, const std::int32_t* idx, std::int64_t n, int bs,
                 double* out)
{
  if (n == 0)
    return;
  pack_kernel<<<static_cast<int>((n * bs + 255) / 256), 256, 0, c->stream>>>(v, idx, n, bs, out);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_unpack(ptb_ctx* c, const double* in, const std::int32_t* idx, std::int64_t n, int bs,
                   double* v)
{
  if (n == 0)
    return;
  unpack_kernel<<<static_cast<int>((n * bs + 255) / 256), 256, 0, c->stream>>>(in, idx, n, bs, v);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_sqnorm(ptb_ctx* c, const double* v, std::int64_t n, double* out_dev)
{
  sqnorm_kernel<<<vec_grid(c, sqnorm_kernel, 2 * n, 9), VEC_THREADS, 0, c->stream>>>(n, v, out_dev, c->partials.p,
                                                               c->tickets.p + 2, peer_view(c),
                                                               next_red_epoch(c));
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

#else
} // namespace
#endif // PTB_HOST_EMU

} // namespace ptb
