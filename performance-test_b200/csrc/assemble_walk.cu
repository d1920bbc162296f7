// P1 matrix assembly along the *star walk* (layout.h build_walk) -- second generation of the
// row-owner gather of assemble.cu, same results to rounding, same reference region
// (fem::assemble_matrix + set_diagonal, poisson_problem.cpp:129-137).
//
// Why: ncu on assemble_matrix_p1<1> (profiles/r01_ncu_full_poisson_assemble_spmv.csv) shows the
// shared-memory pipe as the busiest unit -- per (row, cell) the kernel reads three edge vectors
// (9 LDS.64) and read-modify-writes three accumulators (3 LDS.64 + 3 STS.64): 30 wavefronts. The
// cells of a vertex star can be ordered so that consecutive cells share a face; then two of the
// three edge vectors, one of the three cofactor vectors and two of the three accumulators carry
// over in registers:
//   step      one word per (row, step): which register positions are reloaded, and from where
//   reload    position p: flush a_p into the shared-memory accumulator of its old column, load
//             the new edge vector, a_p = 0                 (1 RMW + 3 LDS.64 instead of 3 RMW + 9)
//   geometry  n_p = e_{p+1} x e_{p+2} is recomputed only if one of its factors changed
//   element   det = e_0 . n_0, c_own = -(n_0 + n_1 + n_2), a_p += c_own . n_p / (6 |det|)
// The step is branch-free (predicated reloads), so the unrolled chunk of steps schedules as one
// block. Shared memory is private per lane (column `lane` of E and acc): no barrier anywhere.
// Default for scalar P1 (measured x1.48 against assemble_matrix_p1<1>, profiles/r01_walk_*);
// PTB_ASM_WALK=0 selects the older kernel.
#include "geom.cuh"
#include "envopt.h"
#include "kernels.h"
#include <climits>
#include <cstdlib>

namespace ptb
{
namespace
{

// Slices per CTA = template parameter WARPS (15.4 KB of shared memory per slice at w = 15).
// Measured at 4 M DOFs (ms, prefetch off/on): 1 warp 0.613/0.599, 2 warps 0.660/0.642,
// 4 warps 0.699/0.687, 7 warps 0.702/0.684 -> one-warp CTAs with the L2 prefetch are the default.
constexpr int WALK_CHUNK = 8; // step words in flight per thread
// L2 prefetch distance in slices: about two generations of resident warps (148 SMs x 14 warps).
#ifdef PTB_HOST_EMU
constexpr int WALK_PF_DIST = 3; // tests/emu: tiny meshes must reach the prefetch address arithmetic
#else
constexpr int WALK_PF_DIST = 4096;
#endif

// Register positions of the walk (three non-owner vertices of the current cell).
struct WalkState
{
  Vec3 e0, e1, e2; // edge vectors owner -> vertex
  Vec3 n0, n1, n2; // n_p = e_{p+1} x e_{p+2}  (= det * grad phi_p)
  double a0, a1, a2, dg;
  std::uint32_t prev; // previous step word (bytes 0..2: columns the positions accumulate into)
};

// EXACT: cofactor vectors by cross_rn (no FMA contraction), so that entries that vanish
// analytically on the lattice are exact zeros -- selected when the operator is going to be
// compacted (PTB_SPMV_COMPACT=1) or by PTB_ASM_EXACT_ZEROS=1; three more FP64 instructions per
// recomputed cofactor vector.
template <int WARPS, bool PREFETCH, bool EXACT = false>
__global__ void __launch_bounds__(WARPS * 32, 14 / WARPS)
assemble_matrix_p1_walk(MatrixArgs A, const std::uint32_t* __restrict__ walk)
{
  constexpr int WALK_WARPS = WARPS;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const std::int32_t slice = blockIdx.x * WALK_WARPS + warp;
  if (slice >= A.n_slices)
    return;
  const std::int64_t mo = A.mat_off[slice], ao = A.adj_off[slice];
  const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
  const int wa = static_cast<int>((A.adj_off[slice + 1] - ao) >> 5);
  const std::int32_t row = slice * 32 + lane;
  const bool live = row < A.n_rows;

  double* E = smem + warp * (A.max_w * 4 * 32) + lane; // E[(k*3+d)*32], private column `lane`
  double* acc = E + A.max_w * 3 * 32;                  // acc[k*32]

  // ---- prologue: step words of the first chunk, row scalars, then the star ------------------
  const std::uint32_t* wp = walk + ao + lane;
  std::uint32_t wd[WALK_CHUNK];
#pragma unroll
  for (int j = 0; j < WALK_CHUNK; ++j)
    wd[j] = j < wa ? __ldg(wp + j * 32) : ADJ_INVALID_DEV;
  const int len = live ? static_cast<int>(A.rowptr[row + 1] - A.rowptr[row]) : 0;
  const bool bc_row = live && A.bc[row];
  const Vec3 X0 = live ? load_point(A.xdof, row) : Vec3{0.0, 0.0, 0.0};

  if constexpr (PREFETCH)
  {
    // Pull the streams of a slice two warp generations ahead into L2: its column indices and
    // step words (one 128-byte line per lane), row pointers and coordinates. The widths of that
    // slice are taken to be this slice's (a hint only: wrong guesses cost nothing but bandwidth).
    const std::int32_t s2 = slice + WALK_PF_DIST;
    if (s2 < A.n_slices)
    {
      const std::int64_t mo2 = A.mat_off[s2], ao2 = A.adj_off[s2];
      // never past the end of an array: clamp to that slice's own widths and to n_rows
      const int w2 = static_cast<int>((A.mat_off[s2 + 1] - mo2) >> 5);
      const int wa2 = static_cast<int>((A.adj_off[s2 + 1] - ao2) >> 5);
      const std::int64_t r2 = static_cast<std::int64_t>(s2) * 32;
      if (lane < w2)
        prefetch_l2(A.cols + mo2 + lane * 32);
      if (lane < wa2)
        prefetch_l2(walk + ao2 + lane * 32);
      if (lane < 8 && r2 + lane * 4 < A.n_rows)
        prefetch_l2(A.xdof + (r2 + lane * 4) * 4);
      if (lane < 2 && r2 + lane * 16 < A.n_rows)
        prefetch_l2(A.rowptr + r2 + lane * 16);
    }
  }

  std::uint32_t bcmask = 0; // bit k: column k of the row is constrained
  int own = -1;             // position of the diagonal in the row
  for (int k0 = 0; k0 < w; k0 += 2 * WALK_CHUNK)
  {
    std::int32_t c[2 * WALK_CHUNK];
#pragma unroll
    for (int j = 0; j < 2 * WALK_CHUNK; ++j)
      c[j] = k0 + j < w ? __ldg(A.cols + mo + (k0 + j) * 32 + lane) : -1;
#pragma unroll
    for (int h = 0; h < 2; ++h)
    {
      Vec3 x[WALK_CHUNK];
      std::uint8_t b[WALK_CHUNK];
#pragma unroll
      for (int j = 0; j < WALK_CHUNK; ++j)
        if (c[h * WALK_CHUNK + j] >= 0)
        {
          x[j] = load_point(A.xdof, c[h * WALK_CHUNK + j]);
          b[j] = __ldg(A.bc + c[h * WALK_CHUNK + j]);
        }
#pragma unroll
      for (int j = 0; j < WALK_CHUNK; ++j)
        if (c[h * WALK_CHUNK + j] >= 0)
        {
          const int k = k0 + h * WALK_CHUNK + j;
          const Vec3 d = x[j] - X0;
          E[(k * 3 + 0) * 32] = d.x;
          E[(k * 3 + 1) * 32] = d.y;
          E[(k * 3 + 2) * 32] = d.z;
          acc[k * 32] = 0.0;
          bcmask |= b[j] ? 1u << k : 0u;
          own = c[h * WALK_CHUNK + j] == row && k < len ? k : own;
        }
    }
  }

  // ---- walk ---------------------------------------------------------------------------------
  WalkState W;
  W.e0 = W.e1 = W.e2 = W.n0 = W.n1 = W.n2 = Vec3{0.0, 0.0, 0.0};
  W.a0 = W.a1 = W.a2 = W.dg = 0.0;
  W.prev = 0;
  auto step = [&](std::uint32_t word) {
    const bool valid = word != ADJ_INVALID_DEV;
    word = valid ? word : (W.prev & 0x00FFFFFFu); // padding: keep everything, add nothing
    const int o0 = word & 0xFFu, o1 = (word >> 8) & 0xFFu, o2 = (word >> 16) & 0xFFu;
    const int p0 = W.prev & 0xFFu, p1 = (W.prev >> 8) & 0xFFu, p2 = (W.prev >> 16) & 0xFFu;
    const bool l0 = word & (1u << 24), l1 = word & (2u << 24), l2 = word & (4u << 24);
    if (l0)
    {
      acc[p0 * 32] += W.a0;
      W.e0 = Vec3{E[(o0 * 3 + 0) * 32], E[(o0 * 3 + 1) * 32], E[(o0 * 3 + 2) * 32]};
      W.a0 = 0.0;
    }
    if (l1)
    {
      acc[p1 * 32] += W.a1;
      W.e1 = Vec3{E[(o1 * 3 + 0) * 32], E[(o1 * 3 + 1) * 32], E[(o1 * 3 + 2) * 32]};
      W.a1 = 0.0;
    }
    if (l2)
    {
      acc[p2 * 32] += W.a2;
      W.e2 = Vec3{E[(o2 * 3 + 0) * 32], E[(o2 * 3 + 1) * 32], E[(o2 * 3 + 2) * 32]};
      W.a2 = 0.0;
    }
    if constexpr (EXACT)
    {
      if (l1 || l2)
        W.n0 = cross_rn(W.e1, W.e2);
      if (l2 || l0)
        W.n1 = cross_rn(W.e2, W.e0);
      if (l0 || l1)
        W.n2 = cross_rn(W.e0, W.e1);
    }
    else
    {
      if (l1 || l2)
        W.n0 = cross(W.e1, W.e2);
      if (l2 || l0)
        W.n1 = cross(W.e2, W.e0);
      if (l0 || l1)
        W.n2 = cross(W.e0, W.e1);
    }
    const double det = dot(W.e0, W.n0);
    const double r = valid ? rcp_nr1(6.0 * fabs(det)) : 0.0;
    const Vec3 c0 = {-(W.n0.x + W.n1.x + W.n2.x), -(W.n0.y + W.n1.y + W.n2.y),
                     -(W.n0.z + W.n1.z + W.n2.z)};
    W.dg = fma(r, dot(c0, c0), W.dg);
    W.a0 = fma(r, dot(c0, W.n0), W.a0);
    W.a1 = fma(r, dot(c0, W.n1), W.a1);
    W.a2 = fma(r, dot(c0, W.n2), W.a2);
    W.prev = word;
  };
  for (int k0 = 0; k0 < wa; k0 += WALK_CHUNK)
  {
    std::uint32_t nx[WALK_CHUNK]; // next chunk in flight while this one is walked
#pragma unroll
    for (int j = 0; j < WALK_CHUNK; ++j)
      nx[j] = k0 + WALK_CHUNK + j < wa ? __ldg(wp + (k0 + WALK_CHUNK + j) * 32) : ADJ_INVALID_DEV;
#pragma unroll
    for (int j = 0; j < WALK_CHUNK; ++j)
      step(wd[j]);
#pragma unroll
    for (int j = 0; j < WALK_CHUNK; ++j)
      wd[j] = nx[j];
  }
  acc[(W.prev & 0xFFu) * 32] += W.a0;
  acc[((W.prev >> 8) & 0xFFu) * 32] += W.a1;
  acc[((W.prev >> 16) & 0xFFu) * 32] += W.a2;

  // ---- epilogue: BC rows/cols -> 0, BC diagonal -> 1, every stored value written once --------
  double diag = 1.0;
  for (int k = 0; k < w; ++k)
  {
    const bool real = k < len, is_own = k == own;
    double val = is_own ? W.dg : acc[k * 32];
    if (bc_row || ((bcmask >> k) & 1u))
      val = is_own ? 1.0 : 0.0;
    if (!real)
      val = 0.0;
    store_stream(A.vals + mo + k * 32 + lane, val); // evict-first: the value stream must not push the stars out of L2
    diag = is_own ? val : diag;
  }
  if (live)
    A.dinv[row] = 1.0 / diag;
}

// ------------------------------------------------------------------------------------------
// Elasticity (BS = 3) along the same walk. CTA = one slice = three warps; warp a accumulates row a
// of the raw tensor  T_j = sum_cells c_own (x) c_j / (6|det|)  per neighbour j (three accumulators
// per register position), and the material law (Elasticity.py:12-15, 33-39) is applied once per
// stored block in the epilogue:   block[a][b] = mu (delta_ab tr T + T[b][a]) + lambda T[a][b],
// which reads the other two warps' rows of T from shared memory. Per (row, cell) and warp this is
// 40 FP64 instructions instead of ~93 (geometry 27 + q = c_own[a]/(6|det|) + 4 x 3 FMAs).
// Shared memory per slice: E [3w][32] (shared star), T [3][3w][32], DG [9][32] (own block),
// C [w][32] int32 columns with the Dirichlet flag in the top bit.
// Measured on the B200 in round 2 (C3: 3.63 -> 2.26 ms); since then the fallback of the edge-ring kernel
// (assemble_ring.cu, 1.02 ms): PTB_ASM_RING=0, or rows longer than 127 columns.
// ------------------------------------------------------------------------------------------
template <bool PREFETCH>
__global__ void __launch_bounds__(96, 4)
assemble_matrix_p1_walk3(MatrixArgs A, const std::uint32_t* __restrict__ walk)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, a = threadIdx.x >> 5; // a = component row handled by this warp
  const std::int32_t slice = blockIdx.x;
  const std::int64_t mo = A.mat_off[slice], ao = A.adj_off[slice];
  const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
  const int wa = static_cast<int>((A.adj_off[slice + 1] - ao) >> 5);
  const std::int32_t row = slice * 32 + lane;
  const bool live = row < A.n_rows;
  const int mw = A.max_w;

  double* E = smem + lane;                                  // E[(k*3+d)*32]
  double* Tall = smem + mw * 3 * 32 + lane;                 // Tall[(x*mw*3 + k*3 + y)*32] = T_k[x][y]
  double* T = Tall + a * (mw * 3 * 32);                     // this warp's row a
  double* DG = smem + mw * 12 * 32 + lane;                  // DG[(x*3+y)*32]
  std::int32_t* C = reinterpret_cast<std::int32_t*>(smem + mw * 12 * 32 + 9 * 32) + lane; // C[k*32]

  const std::uint32_t* wp = walk + ao + lane;
  std::uint32_t wd[WALK_CHUNK];
#pragma unroll
  for (int j = 0; j < WALK_CHUNK; ++j)
    wd[j] = j < wa ? __ldg(wp + j * 32) : ADJ_INVALID_DEV;
  const int len = live ? static_cast<int>(A.rowptr[row + 1] - A.rowptr[row]) : 0;
  const bool bc_row = live && A.bc[row];
  const Vec3 X0 = live ? load_point(A.xdof, row) : Vec3{0.0, 0.0, 0.0};

  if constexpr (PREFETCH)
  {
    const std::int32_t s2 = slice + WALK_PF_DIST / 3;
    if (s2 < A.n_slices && a == 0)
    {
      const std::int64_t mo2 = A.mat_off[s2], ao2 = A.adj_off[s2];
      const int w2 = static_cast<int>((A.mat_off[s2 + 1] - mo2) >> 5);
      const int wa2 = static_cast<int>((A.adj_off[s2 + 1] - ao2) >> 5);
      const std::int64_t r2 = static_cast<std::int64_t>(s2) * 32;
      if (lane < w2)
        prefetch_l2(A.cols + mo2 + lane * 32);
      if (lane < wa2)
        prefetch_l2(walk + ao2 + lane * 32);
      if (lane < 8 && r2 + lane * 4 < A.n_rows)
        prefetch_l2(A.xdof + (r2 + lane * 4) * 4);
      if (lane < 2 && r2 + lane * 16 < A.n_rows)
        prefetch_l2(A.rowptr + r2 + lane * 16);
    }
  }

  // ---- star: warp a stages columns a, a+3, ...; every warp clears its own row of T -----------
  for (int k0 = a; k0 < w; k0 += 3 * WALK_CHUNK)
  {
    std::int32_t c[WALK_CHUNK];
#pragma unroll
    for (int j = 0; j < WALK_CHUNK; ++j)
      c[j] = k0 + 3 * j < w ? __ldg(A.cols + mo + (k0 + 3 * j) * 32 + lane) : -1;
    Vec3 x[WALK_CHUNK];
    std::uint8_t b[WALK_CHUNK];
#pragma unroll
    for (int j = 0; j < WALK_CHUNK; ++j)
      if (c[j] >= 0)
      {
        x[j] = load_point(A.xdof, c[j]);
        b[j] = __ldg(A.bc + c[j]);
      }
#pragma unroll
    for (int j = 0; j < WALK_CHUNK; ++j)
      if (c[j] >= 0)
      {
        const int k = k0 + 3 * j;
        const Vec3 d = x[j] - X0;
        E[(k * 3 + 0) * 32] = d.x;
        E[(k * 3 + 1) * 32] = d.y;
        E[(k * 3 + 2) * 32] = d.z;
        C[k * 32] = c[j] | (b[j] ? INT32_MIN : 0);
      }
  }
  for (int k = 0; k < 3 * w; ++k)
    T[k * 32] = 0.0;
  __syncthreads();

  // ---- walk ---------------------------------------------------------------------------------
  Vec3 e0{0.0, 0.0, 0.0}, e1 = e0, e2 = e0, n0 = e0, n1 = e0, n2 = e0;
  Vec3 t0 = e0, t1 = e0, t2 = e0, dg = e0; // row a of the tensor accumulators (x, y, z = b)
  std::uint32_t prev = 0;
  auto flush = [&](int slot, Vec3& t) {
    T[(slot * 3 + 0) * 32] += t.x;
    T[(slot * 3 + 1) * 32] += t.y;
    T[(slot * 3 + 2) * 32] += t.z;
    t = Vec3{0.0, 0.0, 0.0};
  };
  auto edge = [&](int o) { return Vec3{E[(o * 3 + 0) * 32], E[(o * 3 + 1) * 32], E[(o * 3 + 2) * 32]}; };
  auto step = [&](std::uint32_t word) {
    const bool valid = word != ADJ_INVALID_DEV;
    word = valid ? word : (prev & 0x00FFFFFFu);
    const bool l0 = word & (1u << 24), l1 = word & (2u << 24), l2 = word & (4u << 24);
    if (l0)
    {
      flush(prev & 0xFFu, t0);
      e0 = edge(word & 0xFFu);
    }
    if (l1)
    {
      flush((prev >> 8) & 0xFFu, t1);
      e1 = edge((word >> 8) & 0xFFu);
    }
    if (l2)
    {
      flush((prev >> 16) & 0xFFu, t2);
      e2 = edge((word >> 16) & 0xFFu);
    }
    if (l1 || l2)
      n0 = cross(e1, e2);
    if (l2 || l0)
      n1 = cross(e2, e0);
    if (l0 || l1)
      n2 = cross(e0, e1);
    const double det = dot(e0, n0);
    const double r = valid ? rcp_nr(6.0 * fabs(det)) : 0.0;
    const Vec3 c0 = {-(n0.x + n1.x + n2.x), -(n0.y + n1.y + n2.y), -(n0.z + n1.z + n2.z)};
    const double q = r * comp(c0, a);
    dg = Vec3{fma(q, c0.x, dg.x), fma(q, c0.y, dg.y), fma(q, c0.z, dg.z)};
    t0 = Vec3{fma(q, n0.x, t0.x), fma(q, n0.y, t0.y), fma(q, n0.z, t0.z)};
    t1 = Vec3{fma(q, n1.x, t1.x), fma(q, n1.y, t1.y), fma(q, n1.z, t1.z)};
    t2 = Vec3{fma(q, n2.x, t2.x), fma(q, n2.y, t2.y), fma(q, n2.z, t2.z)};
    prev = word;
  };
  for (int k0 = 0; k0 < wa; k0 += WALK_CHUNK)
  {
    std::uint32_t nx[WALK_CHUNK];
#pragma unroll
    for (int j = 0; j < WALK_CHUNK; ++j)
      nx[j] = k0 + WALK_CHUNK + j < wa ? __ldg(wp + (k0 + WALK_CHUNK + j) * 32) : ADJ_INVALID_DEV;
#pragma unroll
    for (int j = 0; j < WALK_CHUNK; ++j)
      step(wd[j]);
#pragma unroll
    for (int j = 0; j < WALK_CHUNK; ++j)
      wd[j] = nx[j];
  }
  flush(prev & 0xFFu, t0);
  flush((prev >> 8) & 0xFFu, t1);
  flush((prev >> 16) & 0xFFu, t2);
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
    // entry [x][y] of the block's raw tensor: the own block lives in DG, the others in T of warp x
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

} // namespace

#ifndef PTB_HOST_EMU // launchers: device build only
namespace
{
template <int WARPS, bool PREFETCH>
bool launch_walk(ptb_ctx* c, const MatrixArgs& A)
{
  const std::size_t smem = static_cast<std::size_t>(c->max_w) * 4 * 32 * WARPS * sizeof(double);
  if (smem > 227 * 1024)
    return false;
  // the EXACT variant serves the compacted operator; the default stays the faster kernel
  const bool exact = env_flag("PTB_ASM_EXACT_ZEROS", env_flag("PTB_SPMV_COMPACT", false));
  auto kernel = exact ? assemble_matrix_p1_walk<WARPS, PREFETCH, true> : assemble_matrix_p1_walk<WARPS, PREFETCH, false>;
  PTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  kernel<<<(A.n_slices + WARPS - 1) / WARPS, WARPS * 32, smem, c->stream>>>(A, c->walk.p);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
  return true;
}
} // namespace

bool launch_assemble_matrix_walk(ptb_ctx* c, const MatrixArgs& A)
{
  if (c->order != 1 || c->walk.p == nullptr)
    return false;
  if (c->bs == 3)
  {
    const std::size_t smem
        = (static_cast<std::size_t>(c->max_w) * 12 * 32 + 9 * 32) * sizeof(double)
          + static_cast<std::size_t>(c->max_w) * 32 * sizeof(std::int32_t);
    if (smem > 227 * 1024)
      return false;
    const bool pf = env_int("PTB_WALK_PREFETCH", 1) != 0;
    auto kernel = pf ? assemble_matrix_p1_walk3<true> : assemble_matrix_p1_walk3<false>;
    PTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    kernel<<<A.n_slices, 96, smem, c->stream>>>(A, c->walk.p);
    PTB_CUDA(cudaGetLastError());
    c->launches += 1;
    return true;
  }
  if (c->bs != 1 || c->max_w > 32)
    return false;
  // tuning knobs of the opt-in path (A/B on the GPU): slices per CTA and the L2 prefetch
  const int warps = env_int("PTB_WALK_WARPS", 1);
  const bool pf = env_int("PTB_WALK_PREFETCH", 1) != 0;
  switch (warps)
  {
  case 1: return pf ? launch_walk<1, true>(c, A) : launch_walk<1, false>(c, A);
  case 4: return pf ? launch_walk<4, true>(c, A) : launch_walk<4, false>(c, A);
  case 7: return pf ? launch_walk<7, true>(c, A) : launch_walk<7, false>(c, A);
  default: return pf ? launch_walk<2, true>(c, A) : launch_walk<2, false>(c, A);
  }
}

#endif // PTB_HOST_EMU

} // namespace ptb
