// P1 assembly along the one-vertex-per-step star walk (layout.h SellLayout::walk1) with the new
// vertex gathered straight from the coordinate array a few steps ahead -- third generation of the
// row-owner gather (assemble.cu -> assemble_walk.cu -> here). Same reference regions:
//   matrix  fem::assemble_matrix + set_diagonal      poisson_problem.cpp:129-137
//   vector  fem::assemble_vector + bc->set           poisson_problem.cpp:150-155,
//                                                    elasticity_problem.cpp:221-229
//
// Why: ncu on assemble_matrix_p1_walk (profiles/r01_ncu_full_poisson_assemble_walk_4M.csv) puts
// 39 % of the stall samples into the prologue that stages the row star in shared memory
// (mat_off -> columns -> coordinate gather, ~600 instructions, three dependent memory levels), and
// the staged star (11.5 KB per slice) is what limits the kernel to 3.3 warps per scheduler -- yet
// along the walk every staged edge vector is read only 1.7 times. Here nothing is staged: shared
// memory holds the column list (for the gathers) and, for the matrix, the accumulators; step j
// issues the two 16-byte loads of the vertex that step j + GW_AHEAD brings in, so the gather
// latency is covered by four steps of arithmetic, and the footprint drops to 5.8 KB (matrix) /
// 1.9 KB (vector) per slice.
//
// NOT YET RUN ON A GPU (written after the round's GPU budget was spent): opt-in, PTB_ASM_GWALK=1.
// The step encoding and the arithmetic are pinned on the CPU (tests/test_star_walk.py,
// test_direct_gather_walk_reproduces_the_oracle).
#include "geom.cuh"
#include "envopt.h"
#include "kernels.h"
#include <climits>
#include <cstdlib>

namespace ptb
{
namespace
{

constexpr int GW_CHUNK = 8; // step words in flight per thread
constexpr int GW_AHEAD = 4; // gathers in flight per thread (steps of lookahead); divides GW_CHUNK

// A vertex in flight: the two 16-byte halves of its padded coordinates (+ the source term).
template <int NF>
struct InFlight
{
  double2 xy, z_;
  double f[NF > 0 ? NF : 1];
  std::int32_t cw; // the column word it was gathered through (top bit: flag set by the stager)
};

// Issue the gather of the vertex a later step brings in. Padding / no-load steps fetch offset 0
// (always a valid column) so the code stays branch-free.
template <int NF>
__device__ __forceinline__ InFlight<NF> gw_issue(std::uint32_t word, const std::int32_t* C,
                                                 const double* __restrict__ xdof,
                                                 const double* __restrict__ f, int bs, int a)
{
  const bool loads = word != ADJ_INVALID_DEV && ((word >> 16) & 3u) != 3u;
  const int slot = loads ? static_cast<int>(word & 0xFFu) : 0;
  const std::int32_t cw = C[slot * 32];
  const std::int64_t col = cw & INT32_MAX; // the top bit may carry the Dirichlet flag
  const double2* p = reinterpret_cast<const double2*>(xdof + 4 * col);
  InFlight<NF> v;
  v.cw = cw;
  v.xy = __ldg(p);
  v.z_ = __ldg(p + 1);
  if constexpr (NF > 0)
    v.f[0] = __ldg(f + col * bs + a);
  return v;
}

struct StepBits
{
  bool p0, p1, p2, compute;
  int nw, old;
};
__device__ __forceinline__ StepBits gw_decode(std::uint32_t word)
{
  const bool valid = word != ADJ_INVALID_DEV;
  const unsigned pos = valid ? (word >> 16) & 3u : 3u;
  return {pos == 0u, pos == 1u, pos == 2u, valid && ((word >> 18) & 1u) != 0u,
          static_cast<int>(word & 0xFFu), static_cast<int>((word >> 8) & 0xFFu)};
}

// ------------------------------------------------------------------------------------------
// Matrix, scalar P1. One warp = one slice; no barrier (shared memory is private per lane).
// Shared memory per slice: C [w][32] int32 columns, acc [w][32] doubles.
// ------------------------------------------------------------------------------------------
// EXACT: cofactor vectors without FMA contraction (geom.cuh cross_rn), as in assemble_matrix_p1_walk.
template <int WARPS, bool EXACT = false>
__global__ void __launch_bounds__(WARPS * 32)
assemble_matrix_p1_gwalk(MatrixArgs A, const std::uint32_t* __restrict__ walk1,
                         const std::int64_t* __restrict__ walk1_off)
{
  auto cofactor = [](Vec3 a, Vec3 b) {
    if constexpr (EXACT)
      return cross_rn(a, b);
    else
      return cross(a, b);
  };
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const std::int32_t slice = blockIdx.x * WARPS + warp;
  if (slice >= A.n_slices)
    return;
  const std::int64_t mo = A.mat_off[slice], so = walk1_off[slice];
  const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
  const int w1 = static_cast<int>((walk1_off[slice + 1] - so) >> 5);
  const std::int32_t row = slice * 32 + lane;
  const bool live = row < A.n_rows;
  const int mw = A.max_w;

  // per slice: acc (mw doubles per lane) then C (mw int32 per lane)
  double* acc = smem + static_cast<std::size_t>(warp) * (mw * 32 + (mw * 32 + 1) / 2) + lane;
  std::int32_t* C = reinterpret_cast<std::int32_t*>(acc - lane + mw * 32) + lane;

  // ---- prologue -------------------------------------------------------------------------------
  const std::uint32_t word0 = w1 > 0 ? __ldg(walk1 + so + lane) : ADJ_INVALID_DEV;
  const std::uint32_t* wp = walk1 + so + 32 + lane; // step 1
  const int nsteps = w1 - 1;
  std::uint32_t wd[GW_CHUNK];
#pragma unroll
  for (int j = 0; j < GW_CHUNK; ++j)
    wd[j] = j < nsteps ? __ldg(wp + j * 32) : ADJ_INVALID_DEV;
  const int len = live ? static_cast<int>(A.rowptr[row + 1] - A.rowptr[row]) : 0;
  const bool bc_row = live && A.bc[row];
  const Vec3 X0 = live ? load_point(A.xdof, row) : Vec3{0.0, 0.0, 0.0};

  int own = -1;             // position of the diagonal in the row
  std::uint32_t bcmask = 0; // bit k: column k is constrained (w <= 32 checked by the launcher)
  for (int k0 = 0; k0 < w; k0 += 16)
  {
    std::int32_t c[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      c[j] = k0 + j < w ? __ldg(A.cols + mo + (k0 + j) * 32 + lane) : -1;
    std::uint8_t b[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c[j] >= 0)
      {
        const int k = k0 + j;
        C[k * 32] = c[j];
        acc[k * 32] = 0.0;
        own = c[j] == row && k < len ? k : own;
        b[j] = __ldg(A.bc + c[j]);
      }
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c[j] >= 0)
        bcmask |= b[j] ? 1u << (k0 + j) : 0u;
  }

  // ---- step 0 (three vertices) and the first GW_AHEAD gathers --------------------------------
  const bool valid0 = word0 != ADJ_INVALID_DEV;
  const int o0 = valid0 ? word0 & 0xFFu : 0, o1 = valid0 ? (word0 >> 8) & 0xFFu : 0,
            o2 = valid0 ? (word0 >> 16) & 0xFFu : 0;
  Vec3 e0 = load_point(A.xdof, C[o0 * 32]) - X0;
  Vec3 e1 = load_point(A.xdof, C[o1 * 32]) - X0;
  Vec3 e2 = load_point(A.xdof, C[o2 * 32]) - X0;
  InFlight<0> q[GW_AHEAD];
#pragma unroll
  for (int j = 0; j < GW_AHEAD; ++j)
    q[j] = gw_issue<0>(wd[j], C, A.xdof, nullptr, 1, 0);
  int s0 = o0, s1 = o1, s2 = o2; // offsets the three accumulators belong to
  Vec3 n0 = cofactor(e1, e2), n1 = cofactor(e2, e0), n2 = cofactor(e0, e1);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, dg = 0.0;
  auto cell = [&](bool compute) {
    const double det = dot(e0, n0);
    const double r = compute ? rcp_nr(6.0 * fabs(det)) : 0.0;
    const Vec3 c0 = {-(n0.x + n1.x + n2.x), -(n0.y + n1.y + n2.y), -(n0.z + n1.z + n2.z)};
    dg = fma(r, dot(c0, c0), dg);
    a0 = fma(r, dot(c0, n0), a0);
    a1 = fma(r, dot(c0, n1), a1);
    a2 = fma(r, dot(c0, n2), a2);
  };
  cell(valid0);

  // ---- steps 1 .. : one new vertex per step ----------------------------------------------------
  auto step = [&](std::uint32_t word, const InFlight<0>& v) {
    const StepBits S = gw_decode(word);
    const Vec3 xn = {v.xy.x, v.xy.y, v.z_.x};
    // the subtraction writes the position's registers directly (no moves); one flush per step
    const double aold = S.p0 ? a0 : (S.p1 ? a1 : a2);
    if (S.p0 || S.p1 || S.p2)
      acc[S.old * 32] += aold;
    if (S.p0)
      a0 = 0.0, e0 = xn - X0, s0 = S.nw;
    if (S.p1)
      a1 = 0.0, e1 = xn - X0, s1 = S.nw;
    if (S.p2)
      a2 = 0.0, e2 = xn - X0, s2 = S.nw;
    if (S.p1 || S.p2)
      n0 = cofactor(e1, e2);
    if (S.p2 || S.p0)
      n1 = cofactor(e2, e0);
    if (S.p0 || S.p1)
      n2 = cofactor(e0, e1);
    cell(S.compute);
  };
  for (int k0 = 0; k0 < nsteps; k0 += GW_CHUNK)
  {
    std::uint32_t nx[GW_CHUNK]; // next chunk of step words, in flight while this one is walked
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
      nx[j] = k0 + GW_CHUNK + j < nsteps ? __ldg(wp + (k0 + GW_CHUNK + j) * 32) : ADJ_INVALID_DEV;
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
    {
      const InFlight<0> cur = q[j % GW_AHEAD];
      const std::uint32_t ahead = j + GW_AHEAD < GW_CHUNK ? wd[(j + GW_AHEAD) % GW_CHUNK]
                                                          : nx[(j + GW_AHEAD) % GW_CHUNK];
      q[j % GW_AHEAD] = gw_issue<0>(ahead, C, A.xdof, nullptr, 1, 0);
      step(wd[j], cur);
    }
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
      wd[j] = nx[j];
  }
  acc[s0 * 32] += a0;
  acc[s1 * 32] += a1;
  acc[s2 * 32] += a2;

  // ---- epilogue: BC rows/cols -> 0, BC diagonal -> 1, every stored value written once --------
  double diag = 1.0;
  for (int k = 0; k < w; ++k)
  {
    const bool real = k < len, is_own = k == own;
    double val = is_own ? dg : acc[k * 32];
    if (bc_row || ((bcmask >> k) & 1u))
      val = is_own ? 1.0 : 0.0;
    if (!real)
      val = 0.0;
    A.vals[mo + k * 32 + lane] = val;
    diag = is_own ? val : diag;
  }
  if (live)
    A.dinv[row] = 1.0 / diag;
}

// ------------------------------------------------------------------------------------------
// Matrix, elasticity (BS = 3): the tensor accumulation of assemble_matrix_p1_walk3 (warp a keeps
// row a of T_j = sum c_own (x) c_j / 6|det| per neighbour, material law in the epilogue) on the
// one-vertex-per-step walk with direct gathers. CTA = one slice = three warps.
// Shared memory per slice: T [3][3w][32], DG [9][32], C [w][32] int32 (Dirichlet flag in the top
// bit): 36 KB at w = 15 instead of 50 KB with the staged star.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(96, 5)
assemble_matrix_p1_gwalk3(MatrixArgs A, const std::uint32_t* __restrict__ walk1,
                          const std::int64_t* __restrict__ walk1_off)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, a = threadIdx.x >> 5;
  const std::int32_t slice = blockIdx.x;
  const std::int64_t mo = A.mat_off[slice], so = walk1_off[slice];
  const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
  const int w1 = static_cast<int>((walk1_off[slice + 1] - so) >> 5);
  const std::int32_t row = slice * 32 + lane;
  const bool live = row < A.n_rows;
  const int mw = A.max_w;

  double* Tall = smem + lane;                  // Tall[(x*mw*3 + k*3 + y)*32] = T_k[x][y]
  double* T = Tall + a * (mw * 3 * 32);        // this warp's row a
  double* DG = smem + mw * 9 * 32 + lane;      // DG[(x*3+y)*32]
  std::int32_t* C = reinterpret_cast<std::int32_t*>(smem + mw * 9 * 32 + 9 * 32) + lane; // C[k*32]

  const std::uint32_t word0 = w1 > 0 ? __ldg(walk1 + so + lane) : ADJ_INVALID_DEV;
  const std::uint32_t* wp = walk1 + so + 32 + lane;
  const int nsteps = w1 - 1;
  std::uint32_t wd[GW_CHUNK];
#pragma unroll
  for (int j = 0; j < GW_CHUNK; ++j)
    wd[j] = j < nsteps ? __ldg(wp + j * 32) : ADJ_INVALID_DEV;
  const int len = live ? static_cast<int>(A.rowptr[row + 1] - A.rowptr[row]) : 0;
  const bool bc_row = live && A.bc[row];
  const Vec3 X0 = live ? load_point(A.xdof, row) : Vec3{0.0, 0.0, 0.0};

  // columns (warp a stages k = a, a+3, ...) and this warp's accumulators
  for (int k0 = a; k0 < w; k0 += 3 * GW_CHUNK)
  {
    std::int32_t c[GW_CHUNK];
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
      c[j] = k0 + 3 * j < w ? __ldg(A.cols + mo + (k0 + 3 * j) * 32 + lane) : -1;
    std::uint8_t b[GW_CHUNK];
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
      if (c[j] >= 0)
        b[j] = __ldg(A.bc + c[j]);
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
      if (c[j] >= 0)
        C[(k0 + 3 * j) * 32] = c[j] | (b[j] ? INT32_MIN : 0);
  }
  for (int k = 0; k < 3 * w; ++k)
    T[k * 32] = 0.0;
  __syncthreads();

  const bool valid0 = word0 != ADJ_INVALID_DEV;
  const int o0 = valid0 ? word0 & 0xFFu : 0, o1 = valid0 ? (word0 >> 8) & 0xFFu : 0,
            o2 = valid0 ? (word0 >> 16) & 0xFFu : 0;
  Vec3 e0 = load_point(A.xdof, C[o0 * 32] & INT32_MAX) - X0;
  Vec3 e1 = load_point(A.xdof, C[o1 * 32] & INT32_MAX) - X0;
  Vec3 e2 = load_point(A.xdof, C[o2 * 32] & INT32_MAX) - X0;
  InFlight<0> q[GW_AHEAD];
#pragma unroll
  for (int j = 0; j < GW_AHEAD; ++j)
    q[j] = gw_issue<0>(wd[j], C, A.xdof, nullptr, 1, 0);
  int s0 = o0, s1 = o1, s2 = o2;
  Vec3 n0 = cross(e1, e2), n1 = cross(e2, e0), n2 = cross(e0, e1);
  const Vec3 zero = {0.0, 0.0, 0.0};
  Vec3 t0 = zero, t1 = zero, t2 = zero, dg = zero; // row a of the tensor accumulators
  auto cell = [&](bool compute) {
    const double det = dot(e0, n0);
    const double r = compute ? rcp_nr(6.0 * fabs(det)) : 0.0;
    const Vec3 c0 = {-(n0.x + n1.x + n2.x), -(n0.y + n1.y + n2.y), -(n0.z + n1.z + n2.z)};
    const double qa = r * comp(c0, a);
    dg = Vec3{fma(qa, c0.x, dg.x), fma(qa, c0.y, dg.y), fma(qa, c0.z, dg.z)};
    t0 = Vec3{fma(qa, n0.x, t0.x), fma(qa, n0.y, t0.y), fma(qa, n0.z, t0.z)};
    t1 = Vec3{fma(qa, n1.x, t1.x), fma(qa, n1.y, t1.y), fma(qa, n1.z, t1.z)};
    t2 = Vec3{fma(qa, n2.x, t2.x), fma(qa, n2.y, t2.y), fma(qa, n2.z, t2.z)};
  };
  cell(valid0);
  auto flush = [&](int slot, const Vec3& t) {
    T[(slot * 3 + 0) * 32] += t.x;
    T[(slot * 3 + 1) * 32] += t.y;
    T[(slot * 3 + 2) * 32] += t.z;
  };
  auto step = [&](std::uint32_t word, const InFlight<0>& v) {
    const StepBits S = gw_decode(word);
    const Vec3 xn = {v.xy.x, v.xy.y, v.z_.x};
    if (S.p0)
    {
      flush(S.old, t0);
      t0 = zero, e0 = xn - X0, s0 = S.nw;
    }
    if (S.p1)
    {
      flush(S.old, t1);
      t1 = zero, e1 = xn - X0, s1 = S.nw;
    }
    if (S.p2)
    {
      flush(S.old, t2);
      t2 = zero, e2 = xn - X0, s2 = S.nw;
    }
    if (S.p1 || S.p2)
      n0 = cross(e1, e2);
    if (S.p2 || S.p0)
      n1 = cross(e2, e0);
    if (S.p0 || S.p1)
      n2 = cross(e0, e1);
    cell(S.compute);
  };
  for (int k0 = 0; k0 < nsteps; k0 += GW_CHUNK)
  {
    std::uint32_t nx[GW_CHUNK];
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
      nx[j] = k0 + GW_CHUNK + j < nsteps ? __ldg(wp + (k0 + GW_CHUNK + j) * 32) : ADJ_INVALID_DEV;
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
    {
      const InFlight<0> cur = q[j % GW_AHEAD];
      const std::uint32_t ahead = j + GW_AHEAD < GW_CHUNK ? wd[(j + GW_AHEAD) % GW_CHUNK]
                                                          : nx[(j + GW_AHEAD) % GW_CHUNK];
      q[j % GW_AHEAD] = gw_issue<0>(ahead, C, A.xdof, nullptr, 1, 0);
      step(wd[j], cur);
    }
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
      wd[j] = nx[j];
  }
  flush(s0, t0);
  flush(s1, t1);
  flush(s2, t2);
  DG[(a * 3 + 0) * 32] = dg.x;
  DG[(a * 3 + 1) * 32] = dg.y;
  DG[(a * 3 + 2) * 32] = dg.z;
  __syncthreads();

  // ---- epilogue: material law per stored block, BC rows/cols, one write per value -----------
  constexpr double mu = 1.0e6 / (2.0 * (1.0 + 0.3));                       // Elasticity.py:12-15
  constexpr double lmbda = 1.0e6 * 0.3 / ((1.0 + 0.3) * (1.0 - 2.0 * 0.3));
  double diag = 1.0;
  for (int k = 0; k < w; ++k)
  {
    const std::int32_t cw = C[k * 32];
    const bool real = k < len;
    const bool own = real && (cw & INT32_MAX) == row;
    const bool bc_any = bc_row || (real && cw < 0);
    auto Txy = [&](int x, int y) {
      return own ? DG[(x * 3 + y) * 32] : Tall[(x * mw * 3 + k * 3 + y) * 32];
    };
    const double tr = Txy(0, 0) + Txy(1, 1) + Txy(2, 2);
#pragma unroll
    for (int b = 0; b < 3; ++b)
    {
      double val = mu * ((a == b ? tr : 0.0) + Txy(b, a)) + lmbda * Txy(a, b);
      if (bc_any)
        val = (own && a == b) ? 1.0 : 0.0;
      if (!real)
        val = 0.0;
      A.vals[(mo + k * 32) * 9 + (a * 3 + b) * 32 + lane] = val;
      if (own && a == b)
        diag = val;
    }
  }
  if (live)
    A.dinv[static_cast<std::int64_t>(row) * 3 + a] = 1.0 / diag;
}

// ------------------------------------------------------------------------------------------
// Cell vector, P1, BS = 1 or 3: b[row*BS + a] = sum_cells |det|/120 (sum_j f_j + f_own).
// One warp = (slice, component a); warps are independent (each keeps its own copy of the
// column list: 1.9 KB), no barrier, no accumulators in shared memory.
// ------------------------------------------------------------------------------------------
template <int BS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
assemble_vector_p1_gwalk(VectorArgs A, const std::uint32_t* __restrict__ walk1,
                         const std::int64_t* __restrict__ walk1_off)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const std::int32_t gw = blockIdx.x * WARPS + warp;
  const std::int32_t slice = gw / BS;
  const int a = gw - slice * BS;
  if (slice >= A.n_slices)
    return;
  const std::int64_t mo = A.mat_off[slice], so = walk1_off[slice];
  const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
  const int w1 = static_cast<int>((walk1_off[slice + 1] - so) >> 5);
  const std::int32_t row = slice * 32 + lane;
  const bool live = row < A.n_rows;
  std::int32_t* C = reinterpret_cast<std::int32_t*>(smem) + static_cast<std::size_t>(warp) * A.max_w * 32 + lane;

  const std::uint32_t word0 = w1 > 0 ? __ldg(walk1 + so + lane) : ADJ_INVALID_DEV;
  const std::uint32_t* wp = walk1 + so + 32 + lane;
  const int nsteps = w1 - 1;
  std::uint32_t wd[GW_CHUNK];
#pragma unroll
  for (int j = 0; j < GW_CHUNK; ++j)
    wd[j] = j < nsteps ? __ldg(wp + j * 32) : ADJ_INVALID_DEV;
  const bool bc_row = live && A.bc[row];
  const double f_own = live ? __ldg(A.f + static_cast<std::int64_t>(row) * BS + a) : 0.0;
  const Vec3 X0 = live ? load_point(A.xdof, row) : Vec3{0.0, 0.0, 0.0};
  for (int k0 = 0; k0 < w; k0 += 16)
  {
    std::int32_t c[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      c[j] = k0 + j < w ? __ldg(A.cols + mo + (k0 + j) * 32 + lane) : -1;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c[j] >= 0)
        C[(k0 + j) * 32] = c[j];
  }

  const bool valid0 = word0 != ADJ_INVALID_DEV;
  const int o0 = valid0 ? word0 & 0xFFu : 0, o1 = valid0 ? (word0 >> 8) & 0xFFu : 0,
            o2 = valid0 ? (word0 >> 16) & 0xFFu : 0;
  const std::int64_t c0 = C[o0 * 32], c1 = C[o1 * 32], c2 = C[o2 * 32];
  Vec3 e0 = load_point(A.xdof, c0) - X0;
  Vec3 e1 = load_point(A.xdof, c1) - X0;
  Vec3 e2 = load_point(A.xdof, c2) - X0;
  double f0 = __ldg(A.f + c0 * BS + a), f1 = __ldg(A.f + c1 * BS + a), f2 = __ldg(A.f + c2 * BS + a);
  InFlight<1> q[GW_AHEAD];
#pragma unroll
  for (int j = 0; j < GW_AHEAD; ++j)
    q[j] = gw_issue<1>(wd[j], C, A.xdof, A.f, BS, a);
  double sum = 0.0;
  auto cell = [&](bool compute) {
    const double det = dot(e0, cross(e1, e2));
    const double wgt = compute ? fabs(det) * (1.0 / 120.0) : 0.0;
    sum = fma(wgt, ((f_own + f0) + (f1 + f2)) + f_own, sum);
  };
  cell(valid0);
  auto step = [&](std::uint32_t word, const InFlight<1>& v) {
    const StepBits S = gw_decode(word);
    const Vec3 xn = {v.xy.x, v.xy.y, v.z_.x};
    if (S.p0)
      e0 = xn - X0, f0 = v.f[0];
    if (S.p1)
      e1 = xn - X0, f1 = v.f[0];
    if (S.p2)
      e2 = xn - X0, f2 = v.f[0];
    cell(S.compute);
  };
  for (int k0 = 0; k0 < nsteps; k0 += GW_CHUNK)
  {
    std::uint32_t nx[GW_CHUNK];
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
      nx[j] = k0 + GW_CHUNK + j < nsteps ? __ldg(wp + (k0 + GW_CHUNK + j) * 32) : ADJ_INVALID_DEV;
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
    {
      const InFlight<1> cur = q[j % GW_AHEAD];
      const std::uint32_t ahead = j + GW_AHEAD < GW_CHUNK ? wd[(j + GW_AHEAD) % GW_CHUNK]
                                                          : nx[(j + GW_AHEAD) % GW_CHUNK];
      q[j % GW_AHEAD] = gw_issue<1>(ahead, C, A.xdof, A.f, BS, a);
      step(wd[j], cur);
    }
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
      wd[j] = nx[j];
  }
  if (live)
    A.b[static_cast<std::int64_t>(row) * BS + a] = bc_row ? 0.0 : sum;
}

// ------------------------------------------------------------------------------------------
// Matrix-free operator, Poisson P1: y = A p without A (the `action` of the reference's cgpoisson
// problem, cgpoisson_problem.cpp:193-230; form M = action(a, un), Poisson.py:33) along the same
// walk: the new vertex brings its coordinates and its entry of p, the three values of p sit in
// registers next to the edge vectors, no accumulator ever leaves the thread. Dirichlet handling as
// in action_p1_poisson (assemble.cu): constrained columns count as zero, constrained rows return
// p. One warp = one slice; its p.y partial goes to py_partials[slice] (fixed-order reduction by
// reduce_partials). Shared memory: the column list with the Dirichlet flag in the top bit.
// ------------------------------------------------------------------------------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 16 / WARPS)
action_p1_gwalk(VectorArgs A, const std::uint32_t* __restrict__ walk1,
                const std::int64_t* __restrict__ walk1_off, const double* __restrict__ p,
                double* __restrict__ y, double* __restrict__ py_partials)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const std::int32_t slice = blockIdx.x * WARPS + warp;
  if (slice >= A.n_slices)
    return;
  const std::int64_t mo = A.mat_off[slice], so = walk1_off[slice];
  const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
  const int w1 = static_cast<int>((walk1_off[slice + 1] - so) >> 5);
  const std::int32_t row = slice * 32 + lane;
  const bool live = row < A.n_rows;
  std::int32_t* C = reinterpret_cast<std::int32_t*>(smem) + static_cast<std::size_t>(warp) * A.max_w * 32 + lane;

  const std::uint32_t word0 = w1 > 0 ? __ldg(walk1 + so + lane) : ADJ_INVALID_DEV;
  const std::uint32_t* wp = walk1 + so + 32 + lane;
  const int nsteps = w1 - 1;
  std::uint32_t wd[GW_CHUNK];
#pragma unroll
  for (int j = 0; j < GW_CHUNK; ++j)
    wd[j] = j < nsteps ? __ldg(wp + j * 32) : ADJ_INVALID_DEV;
  const bool bc_row = live && A.bc[row];
  const double p_own = live ? __ldg(p + row) : 0.0;
  const Vec3 X0 = live ? load_point(A.xdof, row) : Vec3{0.0, 0.0, 0.0};
  for (int k0 = 0; k0 < w; k0 += 16)
  {
    std::int32_t c[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      c[j] = k0 + j < w ? __ldg(A.cols + mo + (k0 + j) * 32 + lane) : -1;
    std::uint8_t b[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c[j] >= 0)
        b[j] = __ldg(A.bc + c[j]);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c[j] >= 0)
        C[(k0 + j) * 32] = c[j] | (b[j] ? INT32_MIN : 0);
  }

  const bool valid0 = word0 != ADJ_INVALID_DEV;
  const int o0 = valid0 ? word0 & 0xFFu : 0, o1 = valid0 ? (word0 >> 8) & 0xFFu : 0,
            o2 = valid0 ? (word0 >> 16) & 0xFFu : 0;
  const std::int32_t w0c = C[o0 * 32], w1c = C[o1 * 32], w2c = C[o2 * 32];
  Vec3 e0 = load_point(A.xdof, w0c & INT32_MAX) - X0;
  Vec3 e1 = load_point(A.xdof, w1c & INT32_MAX) - X0;
  Vec3 e2 = load_point(A.xdof, w2c & INT32_MAX) - X0;
  double p0 = w0c < 0 ? 0.0 : __ldg(p + (w0c & INT32_MAX));
  double p1 = w1c < 0 ? 0.0 : __ldg(p + (w1c & INT32_MAX));
  double p2 = w2c < 0 ? 0.0 : __ldg(p + (w2c & INT32_MAX));
  InFlight<1> q[GW_AHEAD];
#pragma unroll
  for (int j = 0; j < GW_AHEAD; ++j)
    q[j] = gw_issue<1>(wd[j], C, A.xdof, p, 1, 0);
  Vec3 n0 = cross(e1, e2), n1 = cross(e2, e0), n2 = cross(e0, e1);
  double sum = 0.0;
  auto cell = [&](bool compute) {
    const double det = dot(e0, n0);
    const double r = compute ? rcp_nr(6.0 * fabs(det)) : 0.0;
    const Vec3 c0 = {-(n0.x + n1.x + n2.x), -(n0.y + n1.y + n2.y), -(n0.z + n1.z + n2.z)};
    sum = fma(r, ((dot(c0, c0) * p_own + dot(c0, n0) * p0) + dot(c0, n1) * p1) + dot(c0, n2) * p2, sum);
  };
  cell(valid0);
  auto step = [&](std::uint32_t word, const InFlight<1>& v) {
    const StepBits S = gw_decode(word);
    const Vec3 xn = {v.xy.x, v.xy.y, v.z_.x};
    const double pn = v.cw < 0 ? 0.0 : v.f[0];
    if (S.p0)
      e0 = xn - X0, p0 = pn;
    if (S.p1)
      e1 = xn - X0, p1 = pn;
    if (S.p2)
      e2 = xn - X0, p2 = pn;
    if (S.p1 || S.p2)
      n0 = cross(e1, e2);
    if (S.p2 || S.p0)
      n1 = cross(e2, e0);
    if (S.p0 || S.p1)
      n2 = cross(e0, e1);
    cell(S.compute);
  };
  for (int k0 = 0; k0 < nsteps; k0 += GW_CHUNK)
  {
    std::uint32_t nx[GW_CHUNK];
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
      nx[j] = k0 + GW_CHUNK + j < nsteps ? __ldg(wp + (k0 + GW_CHUNK + j) * 32) : ADJ_INVALID_DEV;
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
    {
      const InFlight<1> cur = q[j % GW_AHEAD];
      const std::uint32_t ahead = j + GW_AHEAD < GW_CHUNK ? wd[(j + GW_AHEAD) % GW_CHUNK]
                                                          : nx[(j + GW_AHEAD) % GW_CHUNK];
      q[j % GW_AHEAD] = gw_issue<1>(ahead, C, A.xdof, p, 1, 0);
      step(wd[j], cur);
    }
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
      wd[j] = nx[j];
  }
  const double yr = bc_row ? p_own : sum;
  double dotv = 0.0;
  if (live)
  {
    y[row] = yr;
    dotv = yr * p_own;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    dotv += __shfl_xor_sync(0xffffffffu, dotv, o);
  if (lane == 0)
    py_partials[slice] = dotv;
}

} // namespace

#ifndef PTB_HOST_EMU // launchers: device build only
namespace
{
template <int WARPS>
void launch_matrix_gwalk(ptb_ctx* c, const MatrixArgs& A)
{
  const std::size_t per_slice = (static_cast<std::size_t>(c->max_w) * 32 + (static_cast<std::size_t>(c->max_w) * 32 + 1) / 2) * sizeof(double);
  const std::size_t smem = per_slice * WARPS;
  const bool exact = env_flag("PTB_ASM_EXACT_ZEROS", env_flag("PTB_SPMV_COMPACT", false));
  auto kernel = exact ? assemble_matrix_p1_gwalk<WARPS, true> : assemble_matrix_p1_gwalk<WARPS, false>;
  PTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kernel<<<(A.n_slices + WARPS - 1) / WARPS, WARPS * 32, smem, c->stream>>>(A, c->walk1.p, c->walk1_off.p);
}

template <int BS, int WARPS>
void launch_vector_gwalk(ptb_ctx* c, const VectorArgs& A)
{
  const std::size_t smem = static_cast<std::size_t>(c->max_w) * 32 * sizeof(std::int32_t) * WARPS;
  auto kernel = assemble_vector_p1_gwalk<BS, WARPS>;
  PTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const std::int64_t warps = static_cast<std::int64_t>(A.n_slices) * BS;
  kernel<<<static_cast<unsigned>((warps + WARPS - 1) / WARPS), WARPS * 32, smem, c->stream>>>(A, c->walk1.p, c->walk1_off.p);
}

} // namespace

bool launch_assemble_matrix_gwalk(ptb_ctx* c, const MatrixArgs& A)
{
  if (c->order != 1 || c->walk1.p == nullptr)
    return false;
  if (c->bs == 3)
  {
    const std::size_t smem = (static_cast<std::size_t>(c->max_w) * 9 * 32 + 9 * 32) * sizeof(double)
                             + static_cast<std::size_t>(c->max_w) * 32 * sizeof(std::int32_t);
    if (smem > 227 * 1024)
      return false;
    PTB_CUDA(cudaFuncSetAttribute(assemble_matrix_p1_gwalk3, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    assemble_matrix_p1_gwalk3<<<A.n_slices, 96, smem, c->stream>>>(A, c->walk1.p, c->walk1_off.p);
    PTB_CUDA(cudaGetLastError());
    c->launches += 1;
    return true;
  }
  if (c->bs != 1 || c->max_w > 32)
    return false;
  switch (env_int("PTB_GWALK_WARPS", 4))
  {
  case 1: launch_matrix_gwalk<1>(c, A); break;
  case 2: launch_matrix_gwalk<2>(c, A); break;
  case 8: launch_matrix_gwalk<8>(c, A); break;
  default: launch_matrix_gwalk<4>(c, A); break;
  }
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
  return true;
}

bool launch_action_gwalk(ptb_ctx* c, const VectorArgs& A, const double* p, double* y, double* py_out)
{
  if (c->order != 1 || c->bs != 1 || c->walk1.p == nullptr)
    return false;
  constexpr int WARPS = 4;
  const std::size_t smem = static_cast<std::size_t>(c->max_w) * 32 * sizeof(std::int32_t) * WARPS;
  PTB_CUDA(cudaFuncSetAttribute(action_p1_gwalk<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  c->mf_partials.alloc(static_cast<std::size_t>(A.n_slices));
  action_p1_gwalk<WARPS><<<(A.n_slices + WARPS - 1) / WARPS, WARPS * 32, smem, c->stream>>>(
      A, c->walk1.p, c->walk1_off.p, p, y, c->mf_partials.p);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
  if (py_out != nullptr)
    launch_reduce_partials(c, A.n_slices, c->mf_partials.p, py_out);
  return true;
}

bool launch_assemble_vector_gwalk(ptb_ctx* c, const VectorArgs& A)
{
  if (c->order != 1 || c->walk1.p == nullptr)
    return false;
  const int warps = env_int("PTB_GWALK_WARPS", 4);
  if (c->bs == 1)
  {
    if (warps == 1)
      launch_vector_gwalk<1, 1>(c, A);
    else if (warps == 8)
      launch_vector_gwalk<1, 8>(c, A);
    else
      launch_vector_gwalk<1, 4>(c, A);
  }
  else
  {
    if (warps == 1)
      launch_vector_gwalk<3, 1>(c, A);
    else if (warps == 8)
      launch_vector_gwalk<3, 8>(c, A);
    else
      launch_vector_gwalk<3, 4>(c, A);
  }
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
  return true;
}

#endif // PTB_HOST_EMU

} // namespace ptb
