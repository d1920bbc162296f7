// Assembly kernels (P1 tetrahedra): the sm_100a replacement of
//   fem::assemble_matrix + MatSetValuesLocal(ADD) + set_diagonal   poisson_problem.cpp:129-137
//   fem::assemble_vector + DirichletBC::set                        poisson_problem.cpp:150-155
// and of the FFCx tabulate_tensor kernels generated from Poisson.py:31-32 / Elasticity.py:39-40.
//
// Scheme ("row-owner gather"): one thread owns one scalar matrix row.
//   prologue  the row's column list *is* its vertex star (P1): the thread stages the edge vectors
//             E[k] = X(col_k) - X(row) of its <= w neighbours in shared memory, once;
//   loop      it walks the row's cells in ascending cell order. A cell is one 32-bit word holding
//             the in-row slot offsets of its four vertices, rotated so the owner comes first (the
//             compressed cell -> CSR-slot map). Three edge vectors come from shared memory, the
//             owner's row of the element matrix is evaluated in registers (cofactor form) and
//             added into per-thread accumulators in shared memory at those offsets;
//   epilogue  BC row/column zeroing, unit BC diagonal (set_diagonal), one coalesced write of every
//             stored value, and 1/diag for Jacobi.
// Each CSR value is written exactly once, by its owner, in a fixed summation order: deterministic,
// no atomics, no zero-fill pass, independent of the partition. All per-row streams are SELL-32 so
// every warp access is a full 128/256-byte line; the cell loop touches global memory once per
// cell (4 bytes per thread).
#include "geom.cuh"
#include "kernels.h"
#include <climits>

namespace ptb
{
namespace
{

constexpr int MAT_THREADS_1 = 128; // 4 slices of scalar rows
constexpr int MAT_THREADS_3 = 192; // 2 slices x 3 components

// Shared-memory staging of the row star. E is [w][3] per row, column-major over the 32 lanes of
// the slice: E[(k*3 + d)*32 + lane].
__device__ __forceinline__ Vec3 star_edge(const double* E, int k, int lane)
{
  return {E[(k * 3 + 0) * 32 + lane], E[(k * 3 + 1) * 32 + lane], E[(k * 3 + 2) * 32 + lane]};
}

// ------------------------------------------------------------------------------------------
// Matrix, P1. BS = 1: Poisson, one warp per slice. BS = 3: elasticity, three warps per slice hold
// the components a = 0, 1, 2 of the 32 block rows and share the staged star.
// Shared memory per slice: E [w*3][32] doubles, C [w][32] int32 columns, then per warp
// acc [w*BS][32] doubles. Global loads are issued in chunks (LD_CHUNK independent loads in flight
// per thread) because the occupancy is bounded by the shared-memory footprint.
// ------------------------------------------------------------------------------------------
constexpr int LD_CHUNK = 8;

struct SliceView
{
  std::int32_t row;
  bool live;
  std::int64_t mo, ao;
  int w, wa;
};

template <typename Args>
__device__ __forceinline__ SliceView slice_view(const Args& A, std::int32_t slice, int lane)
{
  SliceView S;
  const bool ok = slice < A.n_slices;
  S.row = slice * 32 + lane;
  S.live = ok && S.row < A.n_rows;
  S.mo = ok ? A.mat_off[slice] : 0;
  S.w = ok ? static_cast<int>((A.mat_off[slice + 1] - S.mo) >> 5) : 0;
  S.ao = ok ? A.adj_off[slice] : 0;
  S.wa = ok ? static_cast<int>((A.adj_off[slice + 1] - S.ao) >> 5) : 0;
  return S;
}

// Stage the star of the slice: columns k = a, a + BS, ... are handled by this warp.
template <int BS, bool WITH_F>
__device__ __forceinline__ void stage_star(const SliceView& S, int a, int lane,
                                           const std::int32_t* __restrict__ cols,
                                           const double* __restrict__ xdof, Vec3 X0, double* E,
                                           std::int32_t* C, const double* __restrict__ f, double* F,
                                           const std::uint8_t* __restrict__ bc = nullptr)
{
  for (int k0 = a; k0 < S.w; k0 += BS * LD_CHUNK)
  {
    std::int32_t c[LD_CHUNK];
#pragma unroll
    for (int j = 0; j < LD_CHUNK; ++j)
    {
      const int k = k0 + BS * j;
      c[j] = k < S.w ? __ldg(cols + S.mo + k * 32 + lane) : -1;
    }
    Vec3 e[LD_CHUNK];
    double fv[LD_CHUNK][WITH_F ? BS : 1];
    std::uint8_t bcv[LD_CHUNK];
#pragma unroll
    for (int j = 0; j < LD_CHUNK; ++j)
      if (c[j] >= 0)
      {
        e[j] = load_point(xdof, c[j]);
        bcv[j] = bc != nullptr ? __ldg(bc + c[j]) : 0;
        if constexpr (WITH_F)
        {
#pragma unroll
          for (int b = 0; b < BS; ++b)
            fv[j][b] = __ldg(f + static_cast<std::int64_t>(c[j]) * BS + b);
        }
      }
#pragma unroll
    for (int j = 0; j < LD_CHUNK; ++j)
      if (c[j] >= 0)
      {
        const int k = k0 + BS * j;
        const Vec3 d = e[j] - X0;
        E[(k * 3 + 0) * 32 + lane] = d.x;
        E[(k * 3 + 1) * 32 + lane] = d.y;
        E[(k * 3 + 2) * 32 + lane] = d.z;
        if (C != nullptr) // column index, Dirichlet flag of the column in the top bit
          C[k * 32 + lane] = c[j] | (bcv[j] ? INT32_MIN : 0);
        if constexpr (WITH_F)
        {
#pragma unroll
          for (int b = 0; b < BS; ++b)
            F[(k * BS + b) * 32 + lane] = fv[j][b];
        }
      }
  }
}

template <int BS>
__global__ void __launch_bounds__(BS == 1 ? MAT_THREADS_1 : MAT_THREADS_3)
assemble_matrix_p1(MatrixArgs A)
{
  extern __shared__ double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slices_per_cta = (blockDim.x >> 5) / BS;
  const int sl = warp / BS;                // slice within the CTA
  const int a = BS == 1 ? 0 : warp % BS;   // component handled by this warp
  const SliceView S = slice_view(A, blockIdx.x * slices_per_cta + sl, lane);
  const std::int32_t row = S.row;

  // per slice: E (3w) + acc (BS*BS*w) doubles + C (w int32 = w/2 doubles)
  const int per_slice = A.max_w * (3 + BS * BS) * 32 + (A.max_w * 32 + 1) / 2;
  double* E = smem + sl * per_slice;
  double* acc = E + A.max_w * 3 * 32 + a * (A.max_w * BS * 32);
  std::int32_t* C = reinterpret_cast<std::int32_t*>(E + A.max_w * (3 + BS * BS) * 32);

  // ---- every global load of the row is issued up front: the cell words (first W0 cells), the
  // row length and BC flag, then the star staging; nothing later in the kernel waits on DRAM ----
  constexpr int W0 = 24;
  std::uint32_t wd0[W0];
#pragma unroll
  for (int j = 0; j < W0; ++j)
    wd0[j] = j < S.wa ? __ldg(A.adjrot + S.ao + j * 32 + lane) : ADJ_INVALID_DEV;
  const std::int64_t len = S.live ? A.rowptr[row + 1] - A.rowptr[row] : 0;
  const bool bc_row = S.live && A.bc[row];
  const Vec3 X0 = S.live ? load_point(A.xdof, row) : Vec3{0.0, 0.0, 0.0};
  stage_star<BS, false>(S, a, lane, A.cols, A.xdof, X0, E, C, nullptr, nullptr, A.bc);
  for (int k = 0; k < S.w * BS; ++k)
    acc[k * 32 + lane] = 0.0;
  if constexpr (BS == 1)
    __syncwarp();
  else
    __syncthreads();

  constexpr double mu = 1.0e6 / (2.0 * (1.0 + 0.3));                       // Elasticity.py:12-15
  constexpr double lmbda = 1.0e6 * 0.3 / ((1.0 + 0.3) * (1.0 - 2.0 * 0.3));

  // ---- cell loop --------------------------------------------------------------------------
  double dg0 = 0.0, dg1 = 0.0, dg2 = 0.0; // owner's own (diagonal) block row, kept in registers
  auto cell = [&](std::uint32_t word) {
    if (word == ADJ_INVALID_DEV)
      return;
    const int o1 = (word >> 8) & 0xffu, o2 = (word >> 16) & 0xffu, o3 = word >> 24;
    const P1Geom G = p1_geometry(star_edge(E, o1, lane), star_edge(E, o2, lane),
                                 star_edge(E, o3, lane));
    const double s = __drcp_rn(6.0 * fabs(G.det));
    if constexpr (BS == 1)
    {
      dg0 += s * dot(G.c0, G.c0);
      acc[o1 * 32 + lane] += s * dot(G.c0, G.c1);
      acc[o2 * 32 + lane] += s * dot(G.c0, G.c2);
      acc[o3 * 32 + lane] += s * dot(G.c0, G.c3);
    }
    else
    {
      // Ae[(0,a),(t,b)] = s [ mu (delta_ab c0.ct + c0[b] ct[a]) + lambda c0[a] ct[b] ]
      const double c0a = comp(G.c0, a);
      const Vec3 ct[4] = {G.c0, G.c1, G.c2, G.c3};
      const int ot[4] = {0, o1, o2, o3};
#pragma unroll
      for (int t = 0; t < 4; ++t)
      {
        const double d = dot(G.c0, ct[t]);
        const double cta = comp(ct[t], a);
        const double v0 = s * (mu * ((a == 0 ? d : 0.0) + G.c0.x * cta) + lmbda * c0a * ct[t].x);
        const double v1 = s * (mu * ((a == 1 ? d : 0.0) + G.c0.y * cta) + lmbda * c0a * ct[t].y);
        const double v2 = s * (mu * ((a == 2 ? d : 0.0) + G.c0.z * cta) + lmbda * c0a * ct[t].z);
        if (t == 0)
          dg0 += v0, dg1 += v1, dg2 += v2;
        else
        {
          acc[(ot[t] * 3 + 0) * 32 + lane] += v0;
          acc[(ot[t] * 3 + 1) * 32 + lane] += v1;
          acc[(ot[t] * 3 + 2) * 32 + lane] += v2;
        }
      }
    }
  };
#pragma unroll
  for (int j = 0; j < W0; ++j)
    cell(wd0[j]);
  for (int k0 = W0; k0 < S.wa; k0 += LD_CHUNK) // rows with more than W0 cells (unstructured meshes)
  {
    std::uint32_t wd[LD_CHUNK];
#pragma unroll
    for (int j = 0; j < LD_CHUNK; ++j)
      wd[j] = k0 + j < S.wa ? __ldg(A.adjrot + S.ao + (k0 + j) * 32 + lane) : ADJ_INVALID_DEV;
#pragma unroll
    for (int j = 0; j < LD_CHUNK; ++j)
      cell(wd[j]);
  }

  // ---- epilogue: BC rows/cols -> 0, BC diagonal -> 1, write values once, coalesced ----------
  double diag = 1.0;
  {
    for (int k = 0; k < S.w; ++k)
    {
      const std::int32_t cw = C[k * 32 + lane];
      const std::int32_t colk = cw & INT32_MAX;
      const bool real = k < len;
      const bool own = real && colk == row;
      const bool bc_any = bc_row || (real && cw < 0);
      if constexpr (BS == 1)
      {
        double val = own ? dg0 : acc[k * 32 + lane];
        if (bc_any)
          val = own ? 1.0 : 0.0;
        if (!real)
          val = 0.0;
        A.vals[S.mo + k * 32 + lane] = val;
        if (own)
          diag = val;
      }
      else
      {
        const double dg[3] = {dg0, dg1, dg2};
#pragma unroll
        for (int b = 0; b < 3; ++b)
        {
          double val = own ? dg[b] : acc[(k * 3 + b) * 32 + lane];
          if (bc_any)
            val = (own && a == b) ? 1.0 : 0.0;
          if (!real)
            val = 0.0;
          A.vals[(S.mo + k * 32) * 9 + (a * 3 + b) * 32 + lane] = val;
          if (own && a == b)
            diag = val;
        }
      }
    }
  }
  if (S.live)
    A.dinv[static_cast<std::int64_t>(row) * BS + a] = 1.0 / diag;
}

// ------------------------------------------------------------------------------------------
// Vector, P1: b[row] = sum_cells |det|/120 (sum_j f_j + f_row); thread = (row, component).
// Shared memory per slice: E [w*3][32], F [w*BS][32] (source term at the star's vertices).
// ------------------------------------------------------------------------------------------
template <int BS>
__global__ void __launch_bounds__(BS == 1 ? MAT_THREADS_1 : MAT_THREADS_3)
assemble_vector_p1(VectorArgs A)
{
  extern __shared__ double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slices_per_cta = (blockDim.x >> 5) / BS;
  const int sl = warp / BS;
  const int a = BS == 1 ? 0 : warp % BS;
  const SliceView S = slice_view(A, blockIdx.x * slices_per_cta + sl, lane);
  const std::int32_t row = S.row;

  const int per_slice = A.max_w * (3 + BS) * 32;
  double* E = smem + sl * per_slice;
  double* F = E + A.max_w * 3 * 32;

  constexpr int W0 = 24; // all cell words of the row are requested before the star is staged
  std::uint32_t wd0[W0];
#pragma unroll
  for (int j = 0; j < W0; ++j)
    wd0[j] = j < S.wa ? __ldg(A.adjrot + S.ao + j * 32 + lane) : ADJ_INVALID_DEV;
  const bool bc_row = S.live && A.bc[row];
  const double f0 = S.live ? __ldg(A.f + static_cast<std::int64_t>(row) * BS + a) : 0.0;
  const Vec3 X0 = S.live ? load_point(A.xdof, row) : Vec3{0.0, 0.0, 0.0};
  stage_star<BS, true>(S, a, lane, A.cols, A.xdof, X0, E, nullptr, A.f, F);
  if constexpr (BS == 1)
    __syncwarp();
  else
    __syncthreads();

  double sum = 0.0;
  auto cell = [&](std::uint32_t word) {
    if (word == ADJ_INVALID_DEV)
      return;
    const int o1 = (word >> 8) & 0xffu, o2 = (word >> 16) & 0xffu, o3 = word >> 24;
    const Vec3 e1 = star_edge(E, o1, lane), e2 = star_edge(E, o2, lane), e3 = star_edge(E, o3, lane);
    const double det = dot(e1, cross(e2, e3));
    const double f1 = F[(o1 * BS + a) * 32 + lane], f2 = F[(o2 * BS + a) * 32 + lane],
                 f3 = F[(o3 * BS + a) * 32 + lane];
    sum += fabs(det) * (1.0 / 120.0) * (((f0 + f1) + (f2 + f3)) + f0);
  };
#pragma unroll
  for (int j = 0; j < W0; ++j)
    cell(wd0[j]);
  for (int k0 = W0; k0 < S.wa; k0 += LD_CHUNK)
  {
    std::uint32_t wd[LD_CHUNK];
#pragma unroll
    for (int j = 0; j < LD_CHUNK; ++j)
      wd[j] = k0 + j < S.wa ? __ldg(A.adjrot + S.ao + (k0 + j) * 32 + lane) : ADJ_INVALID_DEV;
#pragma unroll
    for (int j = 0; j < LD_CHUNK; ++j)
      cell(wd[j]);
  }
  if (S.live)
    A.b[static_cast<std::int64_t>(row) * BS + a] = bc_row ? 0.0 : sum;
}

// ------------------------------------------------------------------------------------------
// Matrix-free operator, Poisson P1: y = A p without A (the `action` of the reference's cgpoisson
// problem, cgpoisson_problem.cpp:193-230, where the operator is assemble_vector of the action form
// M = action(a, un), Poisson.py:33). Same row-owner gather as the assembly: the row's star (edge
// vectors and the values of p at the star's vertices) is staged once, then every cell contributes
//   y_row += 1/(6|det|) * sum_t (c_0 . c_t) p_t
// from registers. Dirichlet handling reproduces the assembled operator exactly: constrained
// columns contribute nothing (bc->set(un, -1*g) with g = 0 / columns zeroed) and constrained rows
// return p (unit diagonal), so assembled and matrix-free CG take the same iterates to rounding.
// Also returns the local p.y partials (fused dot product, cg.h:65).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MAT_THREADS_1)
action_p1_poisson(VectorArgs A, const double* __restrict__ p, double* __restrict__ y,
                  double* __restrict__ py_partials)
{
  extern __shared__ double smem[];
  __shared__ double red[MAT_THREADS_1 / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const SliceView S = slice_view(A, blockIdx.x * (MAT_THREADS_1 / 32) + warp, lane);
  const std::int32_t row = S.row;
  const int per_slice = A.max_w * 4 * 32;
  double* E = smem + warp * per_slice;
  double* F = E + A.max_w * 3 * 32;

  constexpr int W0 = 24;
  std::uint32_t wd0[W0];
#pragma unroll
  for (int j = 0; j < W0; ++j)
    wd0[j] = j < S.wa ? __ldg(A.adjrot + S.ao + j * 32 + lane) : ADJ_INVALID_DEV;
  const bool bc_row = S.live && A.bc[row];
  const double p0 = S.live ? p[row] : 0.0;
  const Vec3 X0 = S.live ? load_point(A.xdof, row) : Vec3{0.0, 0.0, 0.0};
  // stage E and p at the star (p of constrained columns counts as zero)
  for (int k0 = 0; k0 < S.w; k0 += LD_CHUNK)
  {
    std::int32_t c[LD_CHUNK];
#pragma unroll
    for (int j = 0; j < LD_CHUNK; ++j)
      c[j] = k0 + j < S.w ? __ldg(A.cols + S.mo + (k0 + j) * 32 + lane) : -1;
    Vec3 e[LD_CHUNK];
    double pv[LD_CHUNK];
    std::uint8_t bv[LD_CHUNK];
#pragma unroll
    for (int j = 0; j < LD_CHUNK; ++j)
      if (c[j] >= 0)
      {
        e[j] = load_point(A.xdof, c[j]);
        pv[j] = p[c[j]];
        bv[j] = __ldg(A.bc + c[j]);
      }
#pragma unroll
    for (int j = 0; j < LD_CHUNK; ++j)
      if (c[j] >= 0)
      {
        const int k = k0 + j;
        const Vec3 d = e[j] - X0;
        E[(k * 3 + 0) * 32 + lane] = d.x;
        E[(k * 3 + 1) * 32 + lane] = d.y;
        E[(k * 3 + 2) * 32 + lane] = d.z;
        F[k * 32 + lane] = bv[j] ? 0.0 : pv[j];
      }
  }
  __syncwarp();

  double sum = 0.0;
  auto cell = [&](std::uint32_t word) {
    if (word == ADJ_INVALID_DEV)
      return;
    const int o1 = (word >> 8) & 0xffu, o2 = (word >> 16) & 0xffu, o3 = word >> 24;
    const P1Geom G = p1_geometry(star_edge(E, o1, lane), star_edge(E, o2, lane),
                                 star_edge(E, o3, lane));
    const double s = __drcp_rn(6.0 * fabs(G.det));
    sum += s * (((dot(G.c0, G.c0) * p0 + dot(G.c0, G.c1) * F[o1 * 32 + lane])
                 + dot(G.c0, G.c2) * F[o2 * 32 + lane])
                + dot(G.c0, G.c3) * F[o3 * 32 + lane]);
  };
#pragma unroll
  for (int j = 0; j < W0; ++j)
    cell(wd0[j]);
  for (int k0 = W0; k0 < S.wa; k0 += LD_CHUNK)
  {
    std::uint32_t wd[LD_CHUNK];
#pragma unroll
    for (int j = 0; j < LD_CHUNK; ++j)
      wd[j] = k0 + j < S.wa ? __ldg(A.adjrot + S.ao + (k0 + j) * 32 + lane) : ADJ_INVALID_DEV;
#pragma unroll
    for (int j = 0; j < LD_CHUNK; ++j)
      cell(wd[j]);
  }
  const double yr = bc_row ? p0 : sum;
  double dotv = 0.0;
  if (S.live)
  {
    y[row] = yr;
    dotv = yr * p0;
  }
  // per-CTA partial of p.y in a fixed order (reduced by reduce_partials)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    dotv += __shfl_xor_sync(0xffffffffu, dotv, o);
  if (lane == 0)
    red[warp] = dotv;
  __syncthreads();
  if (tid == 0)
  {
    double t = 0.0;
    for (int w = 0; w < MAT_THREADS_1 / 32; ++w)
      t += red[w];
    py_partials[blockIdx.x] = t;
  }
}

// Deterministic sum of per-CTA partials (one CTA, fixed order).
__global__ void __launch_bounds__(256)
reduce_partials(std::int64_t n, const double* __restrict__ partials, double* out)
{
  __shared__ double sh[256];
  double t = 0.0;
  for (std::int64_t i = threadIdx.x; i < n; i += 256)
    t += partials[i];
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1)
  {
    if (threadIdx.x < o)
      sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    *out = sh[0];
}

// Exterior facets, P1 (Poisson.py:32 g*v*ds): thread = boundary row; facet mass = area/12 (1+delta).
__global__ void assemble_facets_p1(FacetArgs A)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n_frows)
    return;
  const std::int32_t row = A.frow_ids[i];
  if (A.bc[row])
    return;
  double sum = 0.0;
  for (int e = A.frow_ptr[i]; e < A.frow_ptr[i + 1]; ++e)
  {
    const std::int32_t cell = A.fent[2 * e], code = A.fent[2 * e + 1];
    const int lf = code >> 2, li = code & 3;
    const int4 v = __ldg(reinterpret_cast<const int4*>(A.x_dofmap) + cell);
    const int4 d = __ldg(reinterpret_cast<const int4*>(A.dofmap) + cell);
    // the two other facet vertices: the locals != lf, != li
    int o0 = -1, o1 = -1;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (t != lf && t != li)
      {
        if (o0 < 0)
          o0 = t;
        else
          o1 = t;
      }
    const Vec3 X0 = load_point(A.xyz, sel4(v, li)), X1 = load_point(A.xyz, sel4(v, o0)),
               X2 = load_point(A.xyz, sel4(v, o1));
    const Vec3 cr = cross(X1 - X0, X2 - X0);
    const double area2 = sqrt(dot(cr, cr)); // 2 * area
    const double g0 = A.g[sel4(d, li)], g1 = A.g[sel4(d, o0)], g2 = A.g[sel4(d, o1)];
    sum += area2 * (1.0 / 24.0) * ((g0 + g1 + g2) + g0);
  }
  A.b[row] += sum;
}

// Inspection: SELL -> CSR value order.
__global__ void sell_to_csr(std::int32_t n_rows, int bs2, const std::int64_t* __restrict__ rowptr,
                            const std::int64_t* __restrict__ mat_off,
                            const double* __restrict__ vals, double* __restrict__ out)
{
  const std::int32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows)
    return;
  const std::int32_t slice = row >> 5, lane = row & 31;
  const std::int64_t mo = mat_off[slice];
  const std::int64_t r0 = rowptr[row], len = rowptr[row + 1] - r0;
  for (std::int64_t k = 0; k < len; ++k)
    for (int e = 0; e < bs2; ++e)
      out[(r0 + k) * bs2 + e] = vals[(mo + k * 32) * bs2 + e * 32 + lane];
}

// Coordinates by dof index (vertex dofs): xdof[d] = xyz[dof_vertex[d]].
__global__ void gather_xdof(std::int64_t n, const std::int32_t* __restrict__ dof_vertex,
                            const double* __restrict__ xyz, double* __restrict__ xdof)
{
  const std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  const std::int32_t v = dof_vertex[i];
  double2 a = {0.0, 0.0}, b = {0.0, 0.0};
  if (v >= 0)
  {
    const double2* p = reinterpret_cast<const double2*>(xyz + 4 * static_cast<std::int64_t>(v));
    a = p[0], b = p[1];
  }
  double2* q = reinterpret_cast<double2*>(xdof + 4 * i);
  q[0] = a, q[1] = b;
}

std::size_t mat_smem_doubles(int max_w, int bs)
{
  return static_cast<std::size_t>(max_w) * (3 + bs * bs) * 32 + (static_cast<std::size_t>(max_w) * 32 + 1) / 2;
}

__global__ void pad_xyz(std::int64_t n, const double* __restrict__ x3, double* __restrict__ x4)
{
  const std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  double2* q = reinterpret_cast<double2*>(x4 + 4 * i);
  q[0] = make_double2(x3[3 * i], x3[3 * i + 1]);
  q[1] = make_double2(x3[3 * i + 2], 0.0);
}

#ifndef PTB_HOST_EMU // host launchers: device build only
template <typename K>
void set_smem(K kernel, std::size_t smem)
{
  if (smem > 227 * 1024)
    throw std::runtime_error("assembly: row too long for the shared-memory star/accumulators");
  PTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
}

} // namespace

void launch_assemble_matrix(ptb_ctx* c, const MatrixArgs& A)
{
  if (c->order != 1)
    return launch_assemble_matrix_pk(c, A);
  if (A.adjrot == nullptr)
    throw std::runtime_error("assemble_matrix: a P1 row has more than 254 columns");
  if (launch_assemble_matrix_ring(c, A) || launch_assemble_matrix_walk(c, A))
    return;
  if (c->bs == 1)
  {
    const int spc = MAT_THREADS_1 / 32;
    const std::size_t smem = mat_smem_doubles(c->max_w, 1) * spc * sizeof(double);
    set_smem(assemble_matrix_p1<1>, smem);
    assemble_matrix_p1<1><<<(A.n_slices + spc - 1) / spc, MAT_THREADS_1, smem, c->stream>>>(A);
  }
  else
  {
    const int spc = MAT_THREADS_3 / 96;
    const std::size_t smem = mat_smem_doubles(c->max_w, 3) * spc * sizeof(double);
    set_smem(assemble_matrix_p1<3>, smem);
    assemble_matrix_p1<3><<<(A.n_slices + spc - 1) / spc, MAT_THREADS_3, smem, c->stream>>>(A);
  }
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_assemble_vector(ptb_ctx* c, const VectorArgs& A, const FacetArgs& F)
{
  if (c->order != 1)
    return launch_assemble_vector_pk(c, A, F);
  if (A.adjrot == nullptr)
    throw std::runtime_error("assemble_vector: a P1 row has more than 254 columns");
  if (launch_assemble_vector_gwalk(c, A))
  {
    if (F.n_frows > 0 && F.g != nullptr)
    {
      assemble_facets_p1<<<(F.n_frows + 127) / 128, 128, 0, c->stream>>>(F);
      PTB_CUDA(cudaGetLastError());
      c->launches += 1;
    }
    return;
  }
  if (c->bs == 1)
  {
    const int spc = MAT_THREADS_1 / 32;
    const std::size_t smem = static_cast<std::size_t>(c->max_w) * 4 * 32 * spc * sizeof(double);
    set_smem(assemble_vector_p1<1>, smem);
    assemble_vector_p1<1><<<(A.n_slices + spc - 1) / spc, MAT_THREADS_1, smem, c->stream>>>(A);
  }
  else
  {
    const int spc = MAT_THREADS_3 / 96;
    const std::size_t smem = static_cast<std::size_t>(c->max_w) * 6 * 32 * spc * sizeof(double);
    set_smem(assemble_vector_p1<3>, smem);
    assemble_vector_p1<3><<<(A.n_slices + spc - 1) / spc, MAT_THREADS_3, smem, c->stream>>>(A);
  }
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
  if (F.n_frows > 0 && F.g != nullptr)
  {
    assemble_facets_p1<<<(F.n_frows + 127) / 128, 128, 0, c->stream>>>(F);
    PTB_CUDA(cudaGetLastError());
    c->launches += 1;
  }
}

void launch_sell_to_csr(ptb_ctx* c, double* out)
{
  const int bs2 = c->bs * c->bs;
  sell_to_csr<<<(c->n_owned + 127) / 128, 128, 0, c->stream>>>(c->n_owned, bs2, c->rowptr.p,
                                                               c->mat_off.p, c->vals.p, out);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_action_matrix_free(ptb_ctx* c, const VectorArgs& A, const double* p, double* y,
                               double* py_out)
{
  if (c->bs != 1)
    throw std::runtime_error("matrix-free operator: built for the scalar Poisson space only");
  if (c->order != 1)
    return launch_action_matrix_free_pk(c, A, p, y, py_out);
  if (A.adjrot == nullptr)
    throw std::runtime_error("matrix-free operator: a P1 row has more than 254 columns");
  const int spc = MAT_THREADS_1 / 32;
  const int grid = (A.n_slices + spc - 1) / spc;
  const std::size_t smem = static_cast<std::size_t>(c->max_w) * 4 * 32 * spc * sizeof(double);
  set_smem(action_p1_poisson, smem);
  c->mf_partials.alloc(static_cast<std::size_t>(grid));
  action_p1_poisson<<<grid, MAT_THREADS_1, smem, c->stream>>>(A, p, y, c->mf_partials.p);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
  if (py_out != nullptr)
    launch_reduce_partials(c, grid, c->mf_partials.p, py_out);
}

void launch_reduce_partials(ptb_ctx* c, std::int64_t n, const double* partials, double* out)
{
  reduce_partials<<<1, 256, 0, c->stream>>>(n, partials, out);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_pad_xyz(ptb_ctx* c)
{
  const std::int64_t n = c->n_vertices;
  pad_xyz<<<static_cast<int>((n + 255) / 256), 256, 0, c->stream>>>(n, c->xyz3.p, c->xyz.p);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_gather_xdof(ptb_ctx* c)
{
  const std::int64_t n = static_cast<std::int64_t>(c->n_owned) + c->n_ghost;
  gather_xdof<<<static_cast<int>((n + 255) / 256), 256, 0, c->stream>>>(n, c->dof_vertex.p,
                                                                        c->xyz.p, c->xdof.p);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

#else
} // namespace
#endif // PTB_HOST_EMU

} // namespace ptb
