// Assembly kernels (P1 tetrahedra): the sm_100a replacement of
//   fem::assemble_matrix + MatSetValuesLocal(ADD) + set_diagonal   poisson_problem.cpp:129-137
//   fem::assemble_vector + DirichletBC::set                        poisson_problem.cpp:150-155
// and of the FFCx tabulate_tensor kernels generated from Poisson.py:31-32 / Elasticity.py:39-40.
//
// Scheme ("row-owner gather"): one thread owns one scalar matrix row. It walks the
// row's cells in ascending cell order (the precomputed dof -> (cell, local index) list), evaluates
// only *its* row of each element matrix in registers, and adds the entries into per-thread
// accumulators in shared memory at the precomputed in-row slot offsets. Each CSR value is written
// exactly once, by its owner, in a fixed summation order: deterministic, no atomics, no zero-fill
// pass, and the BC row/column zeroing, the unit BC diagonal and the Jacobi diagonal are fused into
// the epilogue. All per-row streams (cell lists, slot offsets, column indices, values) are stored
// SELL-32 so every warp access is a full 128/256-byte line.
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
__device__ __forceinline__ double comp(Vec3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

__device__ __forceinline__ Vec3 load_vertex(const double* __restrict__ xyz, int v)
{
  // padded [n][4]: two 16-byte loads
  const double2* p = reinterpret_cast<const double2*>(xyz + 4 * static_cast<std::int64_t>(v));
  const double2 a = __ldg(p), b = __ldg(p + 1);
  return {a.x, a.y, b.x};
}

__device__ __forceinline__ int sel4(int4 v, int i)
{
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// P1 geometry seen from local vertex li: vertices are taken in the rotated order
// (li, li+1, li+2, li+3) mod 4, so "my" basis function is always number 0. Returns the scaled
// gradients c_t = det * grad(phi_t) (cofactor vectors) and det. Ae[0][t] = c_0.c_t / (6 |det|).
struct P1Geom
{
  Vec3 c0, c1, c2, c3;
  double det;
};

__device__ __forceinline__ P1Geom p1_geometry(const double* __restrict__ xyz, int4 v, int li)
{
  const Vec3 X0 = load_vertex(xyz, sel4(v, li));
  const Vec3 X1 = load_vertex(xyz, sel4(v, (li + 1) & 3));
  const Vec3 X2 = load_vertex(xyz, sel4(v, (li + 2) & 3));
  const Vec3 X3 = load_vertex(xyz, sel4(v, (li + 3) & 3));
  const Vec3 e1 = X1 - X0, e2 = X2 - X0, e3 = X3 - X0;
  P1Geom G;
  G.c1 = cross(e2, e3);
  G.c2 = cross(e3, e1);
  G.c3 = cross(e1, e2);
  G.det = dot(e1, G.c1);
  G.c0 = {-(G.c1.x + G.c2.x + G.c3.x), -(G.c1.y + G.c2.y + G.c3.y), -(G.c1.z + G.c2.z + G.c3.z)};
  return G;
}

// In-row slot offset of rotated local column t (nd = 4, one packed word per pair).
__device__ __forceinline__ int slot4(std::uint32_t word, int li, int t, int so_bits)
{
  const int j = (li + t) & 3;
  return so_bits == 8 ? (word >> (8 * j)) & 0xffu : 0; // 16-bit packing handled by slot4w
}

// ------------------------------------------------------------------------------------------
// Matrix, P1. BS = 1: Poisson, thread = row. BS = 3: elasticity, thread = (node row, component a);
// the three warps of a slice hold a = 0, 1, 2.
// ------------------------------------------------------------------------------------------
template <int BS>
__global__ void __launch_bounds__(BS == 1 ? 128 : 192)
assemble_matrix_p1(MatrixArgs A)
{
  extern __shared__ double acc[]; // [w * BS][blockDim.x]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slices_per_cta = (blockDim.x >> 5) / BS;
  const int a = BS == 1 ? 0 : warp % BS;
  const std::int32_t slice = blockIdx.x * slices_per_cta + warp / BS;
  if (slice >= A.n_slices)
    return;
  const std::int32_t row = slice * 32 + lane;
  const bool live = row < A.n_rows;
  const std::int64_t mo = A.mat_off[slice];
  const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
  const std::int64_t ao = A.adj_off[slice];
  const int wa = static_cast<int>((A.adj_off[slice + 1] - ao) >> 5);
  const int nt = blockDim.x;

  for (int k = 0; k < w * BS; ++k)
    acc[k * nt + tid] = 0.0;

  constexpr double mu = 1.0e6 / (2.0 * (1.0 + 0.3));                       // Elasticity.py:12-15
  constexpr double lmbda = 1.0e6 * 0.3 / ((1.0 + 0.3) * (1.0 - 2.0 * 0.3));

  for (int k = 0; k < wa; ++k)
  {
    const std::uint32_t pair = A.adj[ao + k * 32 + lane];
    if (pair == ADJ_INVALID_DEV)
      continue;
    const std::uint32_t sow = A.adjso[(ao + k * 32) * A.so_words + lane];
    const std::uint32_t sow1 = A.so_bits == 16 ? A.adjso[(ao + k * 32) * A.so_words + 32 + lane] : 0u;
    const std::uint32_t cell = pair >> 2;
    const int li = pair & 3;
    const int4 v = __ldg(reinterpret_cast<const int4*>(A.x_dofmap) + cell);
    const P1Geom G = p1_geometry(A.xyz, v, li);
    const double s = 1.0 / (6.0 * fabs(G.det));
    const Vec3 ct[4] = {G.c0, G.c1, G.c2, G.c3};
#pragma unroll
    for (int t = 0; t < 4; ++t)
    {
      const int j = (li + t) & 3;
      const int o = A.so_bits == 8 ? (sow >> (8 * j)) & 0xffu
                                   : ((j < 2 ? sow : sow1) >> (16 * (j & 1))) & 0xffffu;
      if constexpr (BS == 1)
      {
        acc[o * nt + tid] += s * dot(G.c0, ct[t]);
      }
      else
      {
        // Ae[(0,a),(t,b)] = s [ mu (delta_ab c0.ct + c0[b] ct[a]) + lambda c0[a] ct[b] ]
        const double d = dot(G.c0, ct[t]);
        const double c0a = comp(G.c0, a), cta = comp(ct[t], a);
        const double vb[3] = {s * (mu * ((a == 0 ? d : 0.0) + G.c0.x * cta) + lmbda * c0a * ct[t].x),
                              s * (mu * ((a == 1 ? d : 0.0) + G.c0.y * cta) + lmbda * c0a * ct[t].y),
                              s * (mu * ((a == 2 ? d : 0.0) + G.c0.z * cta) + lmbda * c0a * ct[t].z)};
#pragma unroll
        for (int b = 0; b < 3; ++b)
          acc[(o * 3 + b) * nt + tid] += vb[b];
      }
    }
  }

  // Epilogue: BC rows/cols -> 0, BC diagonal -> 1 (set_diagonal), write values once, coalesced.
  const std::int64_t len = live ? A.rowptr[row + 1] - A.rowptr[row] : 0;
  const bool bc_row = live && A.bc[row];
  double diag = 1.0;
  for (int k = 0; k < w; ++k)
  {
    const std::int32_t col = A.cols[mo + k * 32 + lane];
    const bool real = k < len;
    const bool bc_any = bc_row || (real && A.bc[col]);
    if constexpr (BS == 1)
    {
      double val = acc[k * nt + tid];
      if (bc_any)
        val = (col == row) ? 1.0 : 0.0;
      if (!real)
        val = 0.0;
      A.vals[mo + k * 32 + lane] = val;
      if (real && col == row)
        diag = val;
    }
    else
    {
#pragma unroll
      for (int b = 0; b < 3; ++b)
      {
        double val = acc[(k * 3 + b) * nt + tid];
        if (bc_any)
          val = (col == row && a == b) ? 1.0 : 0.0;
        if (!real)
          val = 0.0;
        A.vals[(mo + k * 32) * 9 + (a * 3 + b) * 32 + lane] = val;
        if (real && col == row && a == b)
          diag = val;
      }
    }
  }
  if (live)
    A.dinv[static_cast<std::int64_t>(row) * BS + a] = 1.0 / diag;
}

// ------------------------------------------------------------------------------------------
// Vector, P1: b[row] = sum_cells |det|/120 (sum_j f_j + f_row); thread = (row, component).
// ------------------------------------------------------------------------------------------
template <int BS>
__global__ void __launch_bounds__(BS == 1 ? 128 : 192)
assemble_vector_p1(VectorArgs A)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slices_per_cta = (blockDim.x >> 5) / BS;
  const int a = BS == 1 ? 0 : warp % BS;
  const std::int32_t slice = blockIdx.x * slices_per_cta + warp / BS;
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
    const std::uint32_t cell = pair >> 2;
    const int li = pair & 3;
    const int4 v = __ldg(reinterpret_cast<const int4*>(A.x_dofmap) + cell);
    const int4 d = __ldg(reinterpret_cast<const int4*>(A.dofmap) + cell);
    const P1Geom G = p1_geometry(A.xyz, v, li);
    const double f0 = __ldg(A.f + static_cast<std::int64_t>(sel4(d, li)) * BS + a);
    const double f1 = __ldg(A.f + static_cast<std::int64_t>(sel4(d, (li + 1) & 3)) * BS + a);
    const double f2 = __ldg(A.f + static_cast<std::int64_t>(sel4(d, (li + 2) & 3)) * BS + a);
    const double f3 = __ldg(A.f + static_cast<std::int64_t>(sel4(d, (li + 3) & 3)) * BS + a);
    sum += fabs(G.det) * (1.0 / 120.0) * (((f0 + f1) + (f2 + f3)) + f0);
  }
  A.b[static_cast<std::int64_t>(row) * BS + a] = A.bc[row] ? 0.0 : sum;
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
    int o[2], n = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (t != lf && t != li && n < 2)
        o[n++] = t;
    const Vec3 X0 = load_vertex(A.xyz, sel4(v, li)), X1 = load_vertex(A.xyz, sel4(v, o[0])),
               X2 = load_vertex(A.xyz, sel4(v, o[1]));
    const Vec3 cr = cross(X1 - X0, X2 - X0);
    const double area2 = sqrt(dot(cr, cr)); // 2 * area
    const double g0 = A.g[sel4(d, li)], g1 = A.g[sel4(d, o[0])], g2 = A.g[sel4(d, o[1])];
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

} // namespace

void launch_assemble_matrix(ptb_ctx* c, const MatrixArgs& A)
{
  if (c->order != 1)
    throw std::runtime_error("assemble_matrix: only order 1 kernels are built in this round");
  if (c->bs == 1)
  {
    const int threads = 128, spc = 4;
    const std::size_t smem = static_cast<std::size_t>(c->max_w) * threads * sizeof(double);
    PTB_CUDA(cudaFuncSetAttribute(assemble_matrix_p1<1>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    assemble_matrix_p1<1><<<(A.n_slices + spc - 1) / spc, threads, smem, c->stream>>>(A);
  }
  else
  {
    const int threads = 192, spc = 2;
    const std::size_t smem = static_cast<std::size_t>(c->max_w) * 3 * threads * sizeof(double);
    if (smem > 227 * 1024)
      throw std::runtime_error("assemble_matrix: row too long for the shared-memory accumulators");
    PTB_CUDA(cudaFuncSetAttribute(assemble_matrix_p1<3>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    assemble_matrix_p1<3><<<(A.n_slices + spc - 1) / spc, threads, smem, c->stream>>>(A);
  }
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_assemble_vector(ptb_ctx* c, const VectorArgs& A, const FacetArgs& F)
{
  if (c->order != 1)
    throw std::runtime_error("assemble_vector: only order 1 kernels are built in this round");
  if (c->bs == 1)
    assemble_vector_p1<1><<<(A.n_slices + 3) / 4, 128, 0, c->stream>>>(A);
  else
    assemble_vector_p1<3><<<(A.n_slices + 1) / 2, 192, 0, c->stream>>>(A);
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

} // namespace ptb
