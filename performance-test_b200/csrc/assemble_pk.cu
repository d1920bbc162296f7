// Assembly kernels for the higher-order Lagrange spaces (P2: nd = 10, P3: nd = 20; scalar Poisson,
// BASELINE config 4). Same row-owner gather as the P1 kernels (assemble.cu): one thread owns one
// matrix row, walks the row's cells in ascending order and evaluates only its own row of each
// element tensor, here in FFCx-free "tensor representation" (SURVEY B4):
//
//     Ae[li][j] = sum_{b<=c} G_bc S[bc][li][j],   G = K K^T |detJ|   (6 geometry factors per cell)
//     be[li]    = |detJ| sum_j M[li][j] f_j
//     facet     = |J t1 x J t2| sum_j MF[lf][li][j] g_j
//
// with the exactly integrated reference tensors of element_tables.h staged in shared memory.
// Replaces the generated tabulate_tensor kernels of Poisson.py:31-32 for degree 2 and 3 and the
// DOLFINx insertion loop (poisson_problem.cpp:129-137,150-155).
#include "element_tables.h"
#include <algorithm>
#include "envopt.h"
#include "kernels.h"

namespace ptb
{
namespace
{

struct Vec3
{
  double x, y, z;
};
__device__ __forceinline__ Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ Vec3 cross(Vec3 a, Vec3 b)
{
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ Vec3 load_point(const double* __restrict__ xyz4, std::int64_t v)
{
  const double2* p = reinterpret_cast<const double2*>(xyz4 + 4 * v);
  const double2 a = __ldg(p), b = __ldg(p + 1);
  return {a.x, a.y, b.x};
}

constexpr int PK_THREADS = 128;

// Reference tensors in shared memory: [plane][li][LD] with an odd row stride LD = ND | 1. A lane reads
// row li of a plane for ITS local index, so at a given (plane, j) the 32 lanes of a warp hit up to ND
// different rows; with the natural stride ND = 20 (160 B) those rows fall into four bank groups
// (five-way conflicts on every load of the P3 kernels' inner loop); with 21 doubles the first 16
// rows are conflict free and rows 16..19 pair up with rows 0..3.
template <int ND>
struct PkTab
{
  static constexpr int LD = ND | 1;
  static constexpr int PLANE = ND * LD;
  __device__ static void stage(double* dst, const double* __restrict__ src, int planes)
  {
    for (int i = threadIdx.x; i < planes * ND * ND; i += blockDim.x)
      dst[(i / ND) * LD + (i % ND)] = src[i];
  }
};

// In-row slot offset of local column j from the packed words of a pair.
template <int ND, bool WIDE>
__device__ __forceinline__ int slot_of(const std::uint32_t* w, int j)
{
  if constexpr (WIDE)
    return (w[j >> 1] >> (16 * (j & 1))) & 0xffffu;
  else
    return (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
}

template <int ND, bool WIDE>
__global__ void __launch_bounds__(PK_THREADS)
assemble_matrix_pk(MatrixArgs A, const double* __restrict__ Sg)
{
  constexpr int NW = WIDE ? (ND + 1) / 2 : (ND + 3) / 4;
  extern __shared__ double smem[];
  double* St = smem; // [6][ND][LD]
  PkTab<ND>::stage(St, Sg, 6);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const std::int32_t slice = blockIdx.x * (PK_THREADS / 32) + warp;
  if (slice >= A.n_slices)
    return;
  const std::int32_t row = slice * 32 + lane;
  const bool live = row < A.n_rows;
  const std::int64_t mo = A.mat_off[slice];
  const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
  const std::int64_t ao = A.adj_off[slice];
  const int wa = static_cast<int>((A.adj_off[slice + 1] - ao) >> 5);
  double* acc = smem + 6 * PkTab<ND>::PLANE + warp * (A.max_w * 32);
  for (int k = 0; k < w; ++k)
    acc[k * 32 + lane] = 0.0;
  __syncwarp();

  for (int k = 0; k < wa; ++k)
  {
    const std::uint32_t pair = A.adj[ao + k * 32 + lane];
    if (pair == ADJ_INVALID_DEV)
      continue;
    std::uint32_t words[NW];
#pragma unroll
    for (int q = 0; q < NW; ++q)
      words[q] = A.adjso[(ao + k * 32) * NW + q * 32 + lane];
    const std::uint32_t cell = pair / ND;
    const int li = pair - cell * ND;
    const int4 v = __ldg(reinterpret_cast<const int4*>(A.x_dofmap) + cell);
    const Vec3 X0 = load_point(A.xyz, v.x);
    const Vec3 e1 = load_point(A.xyz, v.y) - X0, e2 = load_point(A.xyz, v.z) - X0,
               e3 = load_point(A.xyz, v.w) - X0;
    // rows of K = J^-1 are c_b / det (c_b = cofactor vectors); G = |det| K K^T = c_b.c_c / |det|
    const Vec3 c1 = cross(e2, e3), c2 = cross(e3, e1), c3 = cross(e1, e2);
    const double inv = 1.0 / fabs(dot(e1, c1));
    const double G00 = dot(c1, c1) * inv, G01 = dot(c1, c2) * inv, G02 = dot(c1, c3) * inv,
                 G11 = dot(c2, c2) * inv, G12 = dot(c2, c3) * inv, G22 = dot(c3, c3) * inv;
    const double* S = St + li * PkTab<ND>::LD;
#pragma unroll
    for (int j = 0; j < ND; ++j)
    {
      const double val = G00 * S[0 * PkTab<ND>::PLANE + j] + G01 * S[1 * PkTab<ND>::PLANE + j]
                         + G02 * S[2 * PkTab<ND>::PLANE + j] + G11 * S[3 * PkTab<ND>::PLANE + j]
                         + G12 * S[4 * PkTab<ND>::PLANE + j] + G22 * S[5 * PkTab<ND>::PLANE + j];
      acc[slot_of<ND, WIDE>(words, j) * 32 + lane] += val;
    }
  }

  const std::int64_t len = live ? A.rowptr[row + 1] - A.rowptr[row] : 0;
  const bool bc_row = live && A.bc[row];
  double diag = 1.0;
  for (int k = 0; k < w; ++k)
  {
    const std::int32_t col = A.cols[mo + k * 32 + lane];
    const bool real = k < len;
    const bool own = real && col == row;
    double val = acc[k * 32 + lane];
    if (bc_row || (real && A.bc[col]))
      val = own ? 1.0 : 0.0;
    if (!real)
      val = 0.0;
    A.vals[mo + k * 32 + lane] = val;
    if (own)
      diag = val;
  }
  if (live)
    A.dinv[row] = 1.0 / diag;
}

// Geometry factors once per cell (the per-cell half of the element kernel): G = |det| K K^T and |det|
// for every local cell, 64 bytes per cell, read by the (row, cell) pairs of the binned matrix kernel
// instead of gathering four vertices and repeating ~60 FP64 instructions nd times per cell. Same
// operations in the same order as the in-kernel geometry, so the assembled values do not change.
__global__ void __launch_bounds__(256)
cell_geometry_pk(std::int64_t n_cells, const std::int32_t* __restrict__ x_dofmap,
                 const double* __restrict__ xyz, double* __restrict__ cell_g)
{
  const std::int64_t cell = static_cast<std::int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (cell >= n_cells)
    return;
  const int4 v = __ldg(reinterpret_cast<const int4*>(x_dofmap) + cell);
  const Vec3 X0 = load_point(xyz, v.x);
  const Vec3 e1 = load_point(xyz, v.y) - X0, e2 = load_point(xyz, v.z) - X0, e3 = load_point(xyz, v.w) - X0;
  const Vec3 c1 = cross(e2, e3), c2 = cross(e3, e1), c3 = cross(e1, e2);
  const double s = fabs(dot(e1, c1));
  const double inv = 1.0 / s;
  double2* out = reinterpret_cast<double2*>(cell_g + 8 * cell);
  out[0] = double2{dot(c1, c1) * inv, dot(c1, c2) * inv};
  out[1] = double2{dot(c1, c3) * inv, dot(c2, c2) * inv};
  out[2] = double2{dot(c2, c3) * inv, dot(c3, c3) * inv};
  out[3] = double2{s, 0.0};
}

// The same kernel over a list of slices whose rows are at most bin_w long: the accumulators are
// sized by the bin, not by the longest row of the matrix. For P3 the vertex rows (175 columns)
// pin the kernel above to ONE CTA of four warps per SM (198 KB of shared memory) although 26 of 27
// rows are edge and face dofs with 20-60 columns; binned, those run at 8-16 warps per SM.
// Default since round 2 (P3 at 2 M DOFs: 4.75 -> 1.47 ms, profiles/r02/assembly_pk_2M.json); PTB_PK_BINS=0
// selects the kernel above.
template <int ND, bool WIDE, bool CELLG = false>
__global__ void __launch_bounds__(PK_THREADS)
assemble_matrix_pk_binned(MatrixArgs A, const double* __restrict__ Sg,
                          const std::int32_t* __restrict__ slice_list, std::int32_t n_list, int bin_w)
{
  constexpr int NW = WIDE ? (ND + 1) / 2 : (ND + 3) / 4;
  extern __shared__ double smem[];
  double* St = smem; // [6][ND][LD]
  PkTab<ND>::stage(St, Sg, 6);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const std::int32_t item = blockIdx.x * (PK_THREADS / 32) + warp;
  if (item >= n_list)
    return;
  const std::int32_t slice = slice_list[item];
  const std::int32_t row = slice * 32 + lane;
  const bool live = row < A.n_rows;
  const std::int64_t mo = A.mat_off[slice];
  const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
  const std::int64_t ao = A.adj_off[slice];
  const int wa = static_cast<int>((A.adj_off[slice + 1] - ao) >> 5);
  double* acc = smem + 6 * PkTab<ND>::PLANE + warp * (bin_w * 32);
  for (int k = 0; k < w; ++k)
    acc[k * 32 + lane] = 0.0;
  __syncwarp();

  // Cell loop, CB cells per trip: the dependent load levels of a cell (pair and slot words -> vertex
  // ids -> coordinates; with CELLG pair and slot words -> the cell's geometry record) are issued for
  // all CB cells before any is consumed.
  constexpr int CB = CELLG ? 4 : 2;
  for (int k0 = 0; k0 < wa; k0 += CB)
  {
    std::uint32_t pair[CB], words[CB][NW];
#pragma unroll
    for (int u = 0; u < CB; ++u)
    {
      const bool in = k0 + u < wa;
      pair[u] = in ? A.adj[ao + (k0 + u) * 32 + lane] : ADJ_INVALID_DEV;
#pragma unroll
      for (int q = 0; q < NW; ++q)
        words[u][q] = in ? A.adjso[(ao + (k0 + u) * 32) * NW + q * 32 + lane] : 0u;
    }
    int li[CB];
    double G[CB][6];
    if constexpr (CELLG)
    {
      double2 g[CB][3];
#pragma unroll
      for (int u = 0; u < CB; ++u)
      {
        const std::uint32_t cell = pair[u] == ADJ_INVALID_DEV ? 0u : pair[u] / ND; // padding reads cell 0
        li[u] = pair[u] == ADJ_INVALID_DEV ? 0 : static_cast<int>(pair[u] - cell * ND);
        const double2* gp = reinterpret_cast<const double2*>(A.cell_g + 8 * static_cast<std::int64_t>(cell));
        g[u][0] = __ldg(gp), g[u][1] = __ldg(gp + 1), g[u][2] = __ldg(gp + 2);
      }
#pragma unroll
      for (int u = 0; u < CB; ++u)
      {
        G[u][0] = g[u][0].x, G[u][1] = g[u][0].y, G[u][2] = g[u][1].x;
        G[u][3] = g[u][1].y, G[u][4] = g[u][2].x, G[u][5] = g[u][2].y;
      }
    }
    else
    {
      int4 v[CB];
#pragma unroll
      for (int u = 0; u < CB; ++u)
      {
        const std::uint32_t cell = pair[u] == ADJ_INVALID_DEV ? 0u : pair[u] / ND; // padding reads cell 0
        li[u] = pair[u] == ADJ_INVALID_DEV ? 0 : static_cast<int>(pair[u] - cell * ND);
        v[u] = __ldg(reinterpret_cast<const int4*>(A.x_dofmap) + cell);
      }
      Vec3 X[CB][4];
#pragma unroll
      for (int u = 0; u < CB; ++u)
      {
        X[u][0] = load_point(A.xyz, v[u].x), X[u][1] = load_point(A.xyz, v[u].y);
        X[u][2] = load_point(A.xyz, v[u].z), X[u][3] = load_point(A.xyz, v[u].w);
      }
#pragma unroll
      for (int u = 0; u < CB; ++u)
      {
        const Vec3 e1 = X[u][1] - X[u][0], e2 = X[u][2] - X[u][0], e3 = X[u][3] - X[u][0];
        // rows of K = J^-1 are c_b / det (c_b = cofactor vectors); G = |det| K K^T = c_b.c_c / |det|
        const Vec3 c1 = cross(e2, e3), c2 = cross(e3, e1), c3 = cross(e1, e2);
        const double inv = 1.0 / fabs(dot(e1, c1));
        G[u][0] = dot(c1, c1) * inv, G[u][1] = dot(c1, c2) * inv, G[u][2] = dot(c1, c3) * inv;
        G[u][3] = dot(c2, c2) * inv, G[u][4] = dot(c2, c3) * inv, G[u][5] = dot(c3, c3) * inv;
      }
    }
#pragma unroll
    for (int u = 0; u < CB; ++u)
    {
      if (pair[u] == ADJ_INVALID_DEV)
        continue;
      const double G00 = G[u][0], G01 = G[u][1], G02 = G[u][2], G11 = G[u][3], G12 = G[u][4], G22 = G[u][5];
      const double* S = St + li[u] * PkTab<ND>::LD;
#pragma unroll
      for (int j = 0; j < ND; ++j)
      {
        const double val = G00 * S[0 * PkTab<ND>::PLANE + j] + G01 * S[1 * PkTab<ND>::PLANE + j]
                           + G02 * S[2 * PkTab<ND>::PLANE + j] + G11 * S[3 * PkTab<ND>::PLANE + j]
                           + G12 * S[4 * PkTab<ND>::PLANE + j] + G22 * S[5 * PkTab<ND>::PLANE + j];
        acc[slot_of<ND, WIDE>(words[u], j) * 32 + lane] += val;
      }
    }
  }

  // Epilogue in chunks of 32 entries: column indices, then their Dirichlet flags, then the stores --
  // two exposed load latencies per chunk (ncu, P3 at 2 M DOFs: the kernel is bound by exposed load
  // latency at 8-16 warps per SM, long-scoreboard 6-8 stall cycles per issued instruction).
  const std::int64_t len = live ? A.rowptr[row + 1] - A.rowptr[row] : 0;
  const bool bc_row = live && A.bc[row];
  double diag = 1.0;
  constexpr int EB = 32;
  for (int k0 = 0; k0 < w; k0 += EB)
  {
    std::int32_t col[EB];
    std::uint8_t flag[EB];
#pragma unroll
    for (int u = 0; u < EB; ++u)
      col[u] = k0 + u < w ? A.cols[mo + (k0 + u) * 32 + lane] : 0;
#pragma unroll
    for (int u = 0; u < EB; ++u)
      flag[u] = A.bc[col[u]];
#pragma unroll
    for (int u = 0; u < EB; ++u)
    {
      const int k = k0 + u;
      if (k >= w)
        break;
      const bool real = k < len;
      const bool own = real && col[u] == row;
      double val = acc[k * 32 + lane];
      if (bc_row || (real && flag[u]))
        val = own ? 1.0 : 0.0;
      if (!real)
        val = 0.0;
      A.vals[mo + k * 32 + lane] = val;
      if (own)
        diag = val;
    }
  }
  if (live)
    A.dinv[row] = 1.0 / diag;
}

// ------------------------------------------------------------------------------------------
// Linear elasticity on P2 / P3 (Elasticity.py:22-43 with degree 2, 3; the reference's CI runs
// `--problem_type elasticity --order 3`, .github/workflows/ccpp.yml:165-181). Blocked element,
// local index 3 i + a. With u_a[c] = K[c][a] (K = J^-1) and the unsymmetrised reference tensors
// KF[cd][i][j] = int d_c phi_i d_d phi_j (element_tables.h):
//     M^{ab}_ij = |det| sum_cd u_a[c] u_b[d] KF[cd][i][j]            ( = int d_a phi_i d_b phi_j )
//     A[(i,a),(j,b)] = mu (delta_ab sum_c M^{cc}_ij + M^{ba}_ij) + lambda M^{ab}_ij
// (SURVEY K5 for general degree; checked against the oracle's quadrature kernel to 1e-15).
// One warp = (slice, component a) as in the P1 elasticity kernels: the thread owns scalar row
// 3 row + a and accumulates the three entries (b = 0, 1, 2) of each stored block; launched per
// row-length bin, bin_w * 768 B of accumulators per warp. sum_c M^{cc} comes from the symmetrised
// tensors S and G = |det| K K^T exactly as in the scalar kernel.
// ------------------------------------------------------------------------------------------
template <int ND, bool WIDE>
__global__ void assemble_matrix_pk3_binned(MatrixArgs A, const double* __restrict__ Sg,
                                           const double* __restrict__ KFg,
                                           const std::int32_t* __restrict__ slice_list,
                                           std::int32_t n_list, int bin_w)
{
  constexpr int NW = WIDE ? (ND + 1) / 2 : (ND + 3) / 4;
  constexpr double mu = 1.0e6 / (2.0 * (1.0 + 0.3));                       // Elasticity.py:12-15
  constexpr double lmbda = 1.0e6 * 0.3 / ((1.0 + 0.3) * (1.0 - 2.0 * 0.3));
  extern __shared__ double smem[];
  double* St = smem;                          // [6][ND][LD]
  double* Kt = smem + 6 * PkTab<ND>::PLANE;   // [9][ND][LD]
  PkTab<ND>::stage(St, Sg, 6);
  PkTab<ND>::stage(Kt, KFg, 9);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const std::int32_t item = blockIdx.x * nwarps + warp;
  if (item >= 3 * n_list)
    return;
  const std::int32_t slice = slice_list[item / 3];
  const int a = item % 3;
  const std::int32_t row = slice * 32 + lane;
  const bool live = row < A.n_rows;
  const std::int64_t mo = A.mat_off[slice];
  const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
  const std::int64_t ao = A.adj_off[slice];
  const int wa = static_cast<int>((A.adj_off[slice + 1] - ao) >> 5);
  double* acc = smem + 15 * PkTab<ND>::PLANE + static_cast<std::size_t>(warp) * bin_w * 96; // [k][b][lane]
  for (int k = 0; k < 3 * w; ++k)
    acc[k * 32 + lane] = 0.0;
  __syncwarp();

  for (int k = 0; k < wa; ++k)
  {
    const std::uint32_t pair = A.adj[ao + k * 32 + lane];
    if (pair == ADJ_INVALID_DEV)
      continue;
    std::uint32_t words[NW];
#pragma unroll
    for (int q = 0; q < NW; ++q)
      words[q] = A.adjso[(ao + k * 32) * NW + q * 32 + lane];
    const std::uint32_t cell = pair / ND;
    const int li = pair - cell * ND;
    const int4 v = __ldg(reinterpret_cast<const int4*>(A.x_dofmap) + cell);
    const Vec3 X0 = load_point(A.xyz, v.x);
    const Vec3 e1 = load_point(A.xyz, v.y) - X0, e2 = load_point(A.xyz, v.z) - X0,
               e3 = load_point(A.xyz, v.w) - X0;
    // rows of K = J^-1 are the cofactor vectors c_c / det
    const Vec3 c1 = cross(e2, e3), c2 = cross(e3, e1), c3 = cross(e1, e2);
    const double det = dot(e1, c1);
    const double idet = 1.0 / det, s = fabs(det), inv = 1.0 / s;
    const double G00 = dot(c1, c1) * inv, G01 = dot(c1, c2) * inv, G02 = dot(c1, c3) * inv,
                 G11 = dot(c2, c2) * inv, G12 = dot(c2, c3) * inv, G22 = dot(c3, c3) * inv;
    // u_b[c] = K[c][b];  ua = u_a,  su[b] = |det| u_b
    const double kx[3] = {c1.x * idet, c2.x * idet, c3.x * idet}; // u_0
    const double ky[3] = {c1.y * idet, c2.y * idet, c3.y * idet}; // u_1
    const double kz[3] = {c1.z * idet, c2.z * idet, c3.z * idet}; // u_2
    const double* ua = a == 0 ? kx : a == 1 ? ky : kz;
    const double ua0 = ua[0], ua1 = ua[1], ua2 = ua[2];
    const double* S = St + li * PkTab<ND>::LD;
    const double* F = Kt + li * PkTab<ND>::LD;
#pragma unroll 2
    for (int j = 0; j < ND; ++j)
    {
      const double kij = G00 * S[0 * PkTab<ND>::PLANE + j] + G01 * S[1 * PkTab<ND>::PLANE + j] + G02 * S[2 * PkTab<ND>::PLANE + j]
                         + G11 * S[3 * PkTab<ND>::PLANE + j] + G12 * S[4 * PkTab<ND>::PLANE + j] + G22 * S[5 * PkTab<ND>::PLANE + j];
      double f[9];
#pragma unroll
      for (int q = 0; q < 9; ++q)
        f[q] = F[q * PkTab<ND>::PLANE + j]; // KF[c][d], q = 3 c + d
      // z_d = sum_c u_a[c] KF[c][d]  (M^{ab} = |det| z . u_b);  t_c = sum_d KF[c][d] u_a[d]  (M^{ba} = |det| u_b . t)
      const double z0 = ua0 * f[0] + ua1 * f[3] + ua2 * f[6], z1 = ua0 * f[1] + ua1 * f[4] + ua2 * f[7],
                   z2 = ua0 * f[2] + ua1 * f[5] + ua2 * f[8];
      const double t0 = f[0] * ua0 + f[1] * ua1 + f[2] * ua2, t1 = f[3] * ua0 + f[4] * ua1 + f[5] * ua2,
                   t2 = f[6] * ua0 + f[7] * ua1 + f[8] * ua2;
      const double mab0 = s * (z0 * kx[0] + z1 * kx[1] + z2 * kx[2]), mba0 = s * (kx[0] * t0 + kx[1] * t1 + kx[2] * t2);
      const double mab1 = s * (z0 * ky[0] + z1 * ky[1] + z2 * ky[2]), mba1 = s * (ky[0] * t0 + ky[1] * t1 + ky[2] * t2);
      const double mab2 = s * (z0 * kz[0] + z1 * kz[1] + z2 * kz[2]), mba2 = s * (kz[0] * t0 + kz[1] * t1 + kz[2] * t2);
      double* dst = acc + slot_of<ND, WIDE>(words, j) * 96 + lane;
      dst[0] += mu * ((a == 0 ? kij : 0.0) + mba0) + lmbda * mab0;
      dst[32] += mu * ((a == 1 ? kij : 0.0) + mba1) + lmbda * mab1;
      dst[64] += mu * ((a == 2 ? kij : 0.0) + mba2) + lmbda * mab2;
    }
  }

  const std::int64_t len = live ? A.rowptr[row + 1] - A.rowptr[row] : 0;
  const bool bc_row = live && A.bc[row];
  double diag = 1.0;
  for (int k = 0; k < w; ++k)
  {
    const std::int32_t col = A.cols[mo + k * 32 + lane];
    const bool real = k < len;
    const bool own = real && col == row;
    const bool dirichlet = bc_row || (real && A.bc[col]);
#pragma unroll
    for (int b = 0; b < 3; ++b)
    {
      double val = acc[(k * 3 + b) * 32 + lane];
      if (dirichlet)
        val = own && a == b ? 1.0 : 0.0;
      if (!real)
        val = 0.0;
      A.vals[(mo + k * 32) * 9 + (3 * a + b) * 32 + lane] = val;
      if (own && a == b)
        diag = val;
    }
  }
  if (live)
    A.dinv[static_cast<std::int64_t>(row) * 3 + a] = 1.0 / diag;
}

// L = f . v dx for the blocked P2 / P3 space (Elasticity.py:40): be[(i,a)] = |det| sum_j M[i][j] f_j[a].
template <int ND>
__global__ void __launch_bounds__(PK_THREADS)
assemble_vector_pk3(VectorArgs A, const double* __restrict__ Mg)
{
  __shared__ double Mt[PkTab<ND>::PLANE];
  PkTab<ND>::stage(Mt, Mg, 1);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const std::int32_t slice = blockIdx.x * (PK_THREADS / 32) + warp;
  if (slice >= A.n_slices)
    return;
  const std::int32_t row = slice * 32 + lane;
  if (row >= A.n_rows)
    return;
  const std::int64_t ao = A.adj_off[slice];
  const int wa = static_cast<int>((A.adj_off[slice + 1] - ao) >> 5);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < wa; ++k)
  {
    const std::uint32_t pair = A.adj[ao + k * 32 + lane];
    if (pair == ADJ_INVALID_DEV)
      break; // lists are front-packed
    const std::uint32_t cell = pair / ND;
    const int li = pair - cell * ND;
    const int4 v = __ldg(reinterpret_cast<const int4*>(A.x_dofmap) + cell);
    const Vec3 X0 = load_point(A.xyz, v.x);
    const Vec3 e1 = load_point(A.xyz, v.y) - X0, e2 = load_point(A.xyz, v.z) - X0,
               e3 = load_point(A.xyz, v.w) - X0;
    const double det = fabs(dot(e1, cross(e2, e3)));
    const std::int32_t* dofs = A.dofmap + static_cast<std::int64_t>(cell) * ND;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int j = 0; j < ND; ++j)
    {
      const double m = Mt[li * PkTab<ND>::LD + j];
      const double* fj = A.f + 3 * static_cast<std::int64_t>(__ldg(dofs + j));
      t0 += m * __ldg(fj), t1 += m * __ldg(fj + 1), t2 += m * __ldg(fj + 2);
    }
    s0 += det * t0, s1 += det * t1, s2 += det * t2;
  }
  const bool bc = A.bc[row];
  double* b = A.b + 3 * static_cast<std::int64_t>(row);
  b[0] = bc ? 0.0 : s0, b[1] = bc ? 0.0 : s1, b[2] = bc ? 0.0 : s2;
}

template <int ND>
__global__ void __launch_bounds__(PK_THREADS)
assemble_vector_pk(VectorArgs A, const double* __restrict__ Mg)
{
  __shared__ double Mt[PkTab<ND>::PLANE];
  PkTab<ND>::stage(Mt, Mg, 1);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const std::int32_t slice = blockIdx.x * (PK_THREADS / 32) + warp;
  if (slice >= A.n_slices)
    return;
  const std::int32_t row = slice * 32 + lane;
  if (row >= A.n_rows)
    return;
  const std::int64_t ao = A.adj_off[slice];
  const int wa = static_cast<int>((A.adj_off[slice + 1] - ao) >> 5);
  double sum = 0.0;
  for (int k = 0; k < wa; ++k)
  {
    const std::uint32_t pair = A.adj[ao + k * 32 + lane];
    if (pair == ADJ_INVALID_DEV)
      break; // lists are front-packed
    const std::uint32_t cell = pair / ND;
    const int li = pair - cell * ND;
    const int4 v = __ldg(reinterpret_cast<const int4*>(A.x_dofmap) + cell);
    const Vec3 X0 = load_point(A.xyz, v.x);
    const Vec3 e1 = load_point(A.xyz, v.y) - X0, e2 = load_point(A.xyz, v.z) - X0,
               e3 = load_point(A.xyz, v.w) - X0;
    const double det = fabs(dot(e1, cross(e2, e3)));
    const std::int32_t* dofs = A.dofmap + static_cast<std::int64_t>(cell) * ND;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < ND; ++j)
      s += Mt[li * PkTab<ND>::LD + j] * __ldg(A.f + __ldg(dofs + j));
    sum += det * s;
  }
  A.b[row] = A.bc[row] ? 0.0 : sum;
}

template <int ND>
__global__ void assemble_facets_pk(FacetArgs A, const double* __restrict__ MFg)
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
    const int lf = code / ND, li = code - lf * ND;
    const std::int32_t* xv = A.x_dofmap + 4 * static_cast<std::int64_t>(cell);
    // facet lf = face opposite local vertex lf, vertices in ascending local order
    const int a0 = lf == 0 ? 1 : 0, a1 = lf <= 1 ? 2 : 1, a2 = lf <= 2 ? 3 : 2;
    const Vec3 P0 = load_point(A.xyz, xv[a0]), P1 = load_point(A.xyz, xv[a1]),
               P2 = load_point(A.xyz, xv[a2]);
    const Vec3 cr = cross(P1 - P0, P2 - P0);
    const double scale = sqrt(dot(cr, cr)); // |J t1 x J t2| = 2 * area
    const std::int32_t* dofs = A.dofmap + static_cast<std::int64_t>(cell) * ND;
    const double* Mrow = MFg + (lf * ND + li) * ND;
    double s = 0.0;
    for (int j = 0; j < ND; ++j)
      s += Mrow[j] * A.g[dofs[j]];
    sum += scale * s;
  }
  A.b[row] += sum;
}

// Matrix-free operator for P2/P3 (the cgpoisson `action`, cgpoisson_problem.cpp:193-230, with the
// form M = action(a, un) of Poisson.py:33): y_row = sum_cells sum_j Ae[li][j] p_j evaluated as
// sum_m G_m (S_m[li][:] . p_cell), i.e. six length-nd dot products with the reference tensors and
// six FMAs with the geometry factors per (row, cell). Constrained columns contribute nothing and
// constrained rows return p (same operator as the assembled one). Per-CTA partials of p.y.
template <int ND>
__global__ void __launch_bounds__(PK_THREADS)
action_pk(VectorArgs A, const double* __restrict__ Sg, const double* __restrict__ p,
          double* __restrict__ y, double* __restrict__ py_partials)
{
  extern __shared__ double St[]; // [6][ND][LD]
  __shared__ double red[PK_THREADS / 32];
  PkTab<ND>::stage(St, Sg, 6);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const std::int32_t slice = blockIdx.x * (PK_THREADS / 32) + warp;
  const std::int32_t row = slice * 32 + lane;
  const bool active = slice < A.n_slices && row < A.n_rows;
  double dotv = 0.0;
  if (active)
  {
    const std::int64_t ao = A.adj_off[slice];
    const int wa = static_cast<int>((A.adj_off[slice + 1] - ao) >> 5);
    double sum = 0.0;
    for (int k = 0; k < wa; ++k)
    {
      const std::uint32_t pair = A.adj[ao + k * 32 + lane];
      if (pair == ADJ_INVALID_DEV)
        break; // lists are front-packed
      const std::uint32_t cell = pair / ND;
      const int li = pair - cell * ND;
      const int4 v = __ldg(reinterpret_cast<const int4*>(A.x_dofmap) + cell);
      const Vec3 X0 = load_point(A.xyz, v.x);
      const Vec3 e1 = load_point(A.xyz, v.y) - X0, e2 = load_point(A.xyz, v.z) - X0,
                 e3 = load_point(A.xyz, v.w) - X0;
      const Vec3 c1 = cross(e2, e3), c2 = cross(e3, e1), c3 = cross(e1, e2);
      const double inv = 1.0 / fabs(dot(e1, c1));
      const double G00 = dot(c1, c1) * inv, G01 = dot(c1, c2) * inv, G02 = dot(c1, c3) * inv,
                   G11 = dot(c2, c2) * inv, G12 = dot(c2, c3) * inv, G22 = dot(c3, c3) * inv;
      const std::int32_t* dofs = A.dofmap + static_cast<std::int64_t>(cell) * ND;
      const double* S = St + li * PkTab<ND>::LD;
      double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0, t5 = 0.0;
#pragma unroll
      for (int j = 0; j < ND; ++j)
      {
        const std::int32_t dj = __ldg(dofs + j);
        const double pj = A.bc[dj] ? 0.0 : p[dj];
        t0 += S[0 * PkTab<ND>::PLANE + j] * pj;
        t1 += S[1 * PkTab<ND>::PLANE + j] * pj;
        t2 += S[2 * PkTab<ND>::PLANE + j] * pj;
        t3 += S[3 * PkTab<ND>::PLANE + j] * pj;
        t4 += S[4 * PkTab<ND>::PLANE + j] * pj;
        t5 += S[5 * PkTab<ND>::PLANE + j] * pj;
      }
      sum += ((G00 * t0 + G01 * t1) + (G02 * t2 + G11 * t3)) + (G12 * t4 + G22 * t5);
    }
    const double p0 = p[row];
    const double yr = A.bc[row] ? p0 : sum;
    y[row] = yr;
    dotv = yr * p0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    dotv += __shfl_xor_sync(0xffffffffu, dotv, o);
  if (lane == 0)
    red[warp] = dotv;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double t = 0.0;
    for (int w = 0; w < PK_THREADS / 32; ++w)
      t += red[w];
    py_partials[blockIdx.x] = t;
  }
}

#ifndef PTB_HOST_EMU // host launchers: device build only
void ensure_tables(ptb_ctx* c)
{
  if (c->tab_order == c->order)
    return;
  const bool p2 = c->order == 2;
  const int nd = c->nd;
  c->tab_S.upload(p2 ? tables::S_P2 : tables::S_P3, 6 * nd * nd, c->stream);
  c->tab_M.upload(p2 ? tables::M_P2 : tables::M_P3, nd * nd, c->stream);
  c->tab_MF.upload(p2 ? tables::MF_P2 : tables::MF_P3, 4 * nd * nd, c->stream);
  c->tab_KF.upload(p2 ? tables::KF_P2 : tables::KF_P3, 9 * nd * nd, c->stream);
  c->tab_order = c->order;
}

// The launches of the row-length classes are independent (disjoint slices): launch(b, stream) of every
// non-empty class goes to its own side stream between a fork and a join event on the context's
// stream, so the classes with few, long rows (P3 vertex rows: 562 CTAs at one CTA per SM, 377 us at
// 2 M DOFs, ncu profiles/r02/ncu_pk_binned_p3_2M.csv) run beside the large ones instead of after them.
// PTB_PK_CONCURRENT=0 queues them on the context's stream one after the other.
template <typename Launch>
void run_bins(ptb_ctx* c, Launch launch)
{
  const std::size_t nb = c->pk_bin_off.empty() ? 0 : c->pk_bin_off.size() - 1;
  if (!env_flag("PTB_PK_CONCURRENT", true))
  {
    for (std::size_t b = 0; b < nb; ++b)
      if (c->pk_bin_off[b + 1] > c->pk_bin_off[b])
        launch(b, c->stream);
    return;
  }
  if (!c->bin_fork)
  {
    PTB_CUDA(cudaEventCreateWithFlags(&c->bin_fork, cudaEventDisableTiming));
    for (int i = 0; i < ptb_ctx::N_BIN_STREAMS; ++i)
    {
      PTB_CUDA(cudaStreamCreateWithFlags(&c->bin_streams[i], cudaStreamNonBlocking));
      PTB_CUDA(cudaEventCreateWithFlags(&c->bin_join[i], cudaEventDisableTiming));
    }
  }
  PTB_CUDA(cudaEventRecord(c->bin_fork, c->stream));
  bool used[ptb_ctx::N_BIN_STREAMS] = {};
  // the classes with the longest rows first: they have the fewest CTAs and the longest tails
  for (std::size_t b = nb; b-- > 0;)
  {
    if (c->pk_bin_off[b + 1] == c->pk_bin_off[b])
      continue;
    const int i = static_cast<int>(b % ptb_ctx::N_BIN_STREAMS);
    if (!used[i])
      PTB_CUDA(cudaStreamWaitEvent(c->bin_streams[i], c->bin_fork, 0));
    used[i] = true;
    launch(b, c->bin_streams[i]);
  }
  for (int i = 0; i < ptb_ctx::N_BIN_STREAMS; ++i)
    if (used[i])
    {
      PTB_CUDA(cudaEventRecord(c->bin_join[i], c->bin_streams[i]));
      PTB_CUDA(cudaStreamWaitEvent(c->stream, c->bin_join[i], 0));
    }
}

template <int ND, bool WIDE>
void launch_matrix_bins(ptb_ctx* c, const MatrixArgs& A)
{
  run_bins(c, [&](std::size_t b, cudaStream_t stream) {
    const std::int32_t n = c->pk_bin_off[b + 1] - c->pk_bin_off[b];
    const int bin_w = c->pk_bin_w[b];
    const std::size_t smem
        = (static_cast<std::size_t>(6) * ND * (ND | 1) + static_cast<std::size_t>(bin_w) * 32 * (PK_THREADS / 32))
          * sizeof(double);
    if (smem > 227 * 1024)
      throw std::runtime_error("assemble_matrix: row too long for the shared-memory accumulators");
    PTB_CUDA(cudaFuncSetAttribute(assemble_matrix_pk_binned<ND, WIDE>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    PTB_CUDA(cudaFuncSetAttribute(assemble_matrix_pk_binned<ND, WIDE, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const int grid = (n + PK_THREADS / 32 - 1) / (PK_THREADS / 32);
    if (A.cell_g != nullptr)
      assemble_matrix_pk_binned<ND, WIDE, true><<<grid, PK_THREADS, smem, stream>>>(
          A, c->tab_S.p, c->pk_bin_slices.p + c->pk_bin_off[b], n, bin_w);
    else
      assemble_matrix_pk_binned<ND, WIDE><<<grid, PK_THREADS, smem, stream>>>(
          A, c->tab_S.p, c->pk_bin_slices.p + c->pk_bin_off[b], n, bin_w);
    PTB_CUDA(cudaGetLastError());
    c->launches += 1;
  });
}

template <int ND, bool WIDE>
void launch_matrix3_bins(ptb_ctx* c, const MatrixArgs& A)
{
  if (c->pk_bin_slices.p == nullptr)
    throw std::runtime_error("assemble_matrix: the elasticity P2/P3 kernel needs the row-length bins");
  PTB_CUDA(cudaFuncSetAttribute(assemble_matrix_pk3_binned<ND, WIDE>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  run_bins(c, [&](std::size_t b, cudaStream_t stream) {
    const std::int32_t n = c->pk_bin_off[b + 1] - c->pk_bin_off[b];
    const int bin_w = c->pk_bin_w[b];
    const std::size_t tables_bytes = static_cast<std::size_t>(15) * ND * (ND | 1) * sizeof(double);
    const std::size_t per_warp = static_cast<std::size_t>(bin_w) * 96 * sizeof(double);
    if (tables_bytes + per_warp > 227 * 1024)
      throw std::runtime_error("assemble_matrix: row too long for the shared-memory accumulators");
    const int warps = static_cast<int>(std::min<std::size_t>(4, (227 * 1024 - tables_bytes) / per_warp));
    const std::int64_t items = static_cast<std::int64_t>(n) * 3;
    const int grid = static_cast<int>((items + warps - 1) / warps);
    assemble_matrix_pk3_binned<ND, WIDE><<<grid, warps * 32, tables_bytes + per_warp * warps, stream>>>(
        A, c->tab_S.p, c->tab_KF.p, c->pk_bin_slices.p + c->pk_bin_off[b], n, bin_w);
    PTB_CUDA(cudaGetLastError());
  });
}

template <int ND>
void launch_matrix(ptb_ctx* c, const MatrixArgs& A)
{
  if (c->bs == 3)
  {
    if (c->so_bits == 8)
      launch_matrix3_bins<ND, false>(c, A);
    else
      launch_matrix3_bins<ND, true>(c, A);
    return;
  }
  if (env_flag("PTB_PK_BINS", true) && c->pk_bin_slices.p != nullptr)
  {
    MatrixArgs B = A;
    // Opt-in: measured on the B200 without gain (profiles/r02/assembly_pk_cellg_ab.txt: P3 1.246 vs
    // 1.255 ms, P2 0.691 vs 0.676 ms at 2 M DOFs) -- the kernel waits on its index streams and its
    // epilogue, not on the geometry chain.
    if (env_flag("PTB_PK_CELLG", false))
    {
      // per-cell half of the element kernel first: geometry factors of every local cell
      c->cell_g.alloc(static_cast<std::size_t>(c->n_cells) * 8);
      cell_geometry_pk<<<static_cast<unsigned>((c->n_cells + 255) / 256), 256, 0, c->stream>>>(
          c->n_cells, A.x_dofmap, A.xyz, c->cell_g.p);
      PTB_CUDA(cudaGetLastError());
      c->launches += 1;
      B.cell_g = c->cell_g.p;
    }
    if (c->so_bits == 8)
      launch_matrix_bins<ND, false>(c, B);
    else
      launch_matrix_bins<ND, true>(c, B);
    c->launches -= 1; // the caller counts one launch
    return;
  }
  const std::size_t smem
      = (static_cast<std::size_t>(6) * ND * (ND | 1) + static_cast<std::size_t>(c->max_w) * 32 * (PK_THREADS / 32))
        * sizeof(double);
  if (smem > 227 * 1024)
    throw std::runtime_error("assemble_matrix: row too long for the shared-memory accumulators");
  const int grid = (A.n_slices + PK_THREADS / 32 - 1) / (PK_THREADS / 32);
  if (c->so_bits == 8)
  {
    PTB_CUDA(cudaFuncSetAttribute(assemble_matrix_pk<ND, false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    assemble_matrix_pk<ND, false><<<grid, PK_THREADS, smem, c->stream>>>(A, c->tab_S.p);
  }
  else
  {
    PTB_CUDA(cudaFuncSetAttribute(assemble_matrix_pk<ND, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    assemble_matrix_pk<ND, true><<<grid, PK_THREADS, smem, c->stream>>>(A, c->tab_S.p);
  }
}

} // namespace

void launch_assemble_matrix_pk(ptb_ctx* c, const MatrixArgs& A)
{
  ensure_tables(c);
  if (c->order == 2)
    launch_matrix<10>(c, A);
  else
    launch_matrix<20>(c, A);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_action_matrix_free_pk(ptb_ctx* c, const VectorArgs& A, const double* p, double* y,
                                  double* py_out)
{
  if (c->bs != 1)
    throw std::runtime_error("matrix-free operator: built for the scalar Poisson space only");
  ensure_tables(c);
  const int grid = (A.n_slices + PK_THREADS / 32 - 1) / (PK_THREADS / 32);
  c->mf_partials.alloc(static_cast<std::size_t>(grid));
  if (c->order == 2)
  {
    const std::size_t smem = 6 * 10 * 11 * sizeof(double);
    action_pk<10><<<grid, PK_THREADS, smem, c->stream>>>(A, c->tab_S.p, p, y, c->mf_partials.p);
  }
  else
  {
    const std::size_t smem = 6 * 20 * 21 * sizeof(double);
    action_pk<20><<<grid, PK_THREADS, smem, c->stream>>>(A, c->tab_S.p, p, y, c->mf_partials.p);
  }
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
  if (py_out != nullptr)
    launch_reduce_partials(c, grid, c->mf_partials.p, py_out);
}

void launch_assemble_vector_pk(ptb_ctx* c, const VectorArgs& A, const FacetArgs& F)
{
  ensure_tables(c);
  const int grid = (A.n_slices + PK_THREADS / 32 - 1) / (PK_THREADS / 32);
  if (c->bs == 3)
  {
    // Elasticity.py:40: cells only, no exterior-facet term
    if (c->order == 2)
      assemble_vector_pk3<10><<<grid, PK_THREADS, 0, c->stream>>>(A, c->tab_M.p);
    else
      assemble_vector_pk3<20><<<grid, PK_THREADS, 0, c->stream>>>(A, c->tab_M.p);
    PTB_CUDA(cudaGetLastError());
    c->launches += 1;
    return;
  }
  if (c->order == 2)
    assemble_vector_pk<10><<<grid, PK_THREADS, 0, c->stream>>>(A, c->tab_M.p);
  else
    assemble_vector_pk<20><<<grid, PK_THREADS, 0, c->stream>>>(A, c->tab_M.p);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
  if (F.n_frows > 0 && F.g != nullptr)
  {
    const int fg = (F.n_frows + 127) / 128;
    if (c->order == 2)
      assemble_facets_pk<10><<<fg, 128, 0, c->stream>>>(F, c->tab_MF.p);
    else
      assemble_facets_pk<20><<<fg, 128, 0, c->stream>>>(F, c->tab_MF.p);
    PTB_CUDA(cudaGetLastError());
    c->launches += 1;
  }
}

#else
} // namespace
#endif // PTB_HOST_EMU

} // namespace ptb
