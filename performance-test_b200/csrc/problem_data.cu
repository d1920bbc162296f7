// Device-side problem data -- SURVEY 8f row 3: the reference's "ZZZ Create boundary conditions" and
// "ZZZ Create RHS function" regions evaluated on the device instead of being passed in
// (ptb_set_bc / ptb_set_source):
//   locate_bc_facets     mesh::locate_entities (facets whose three vertices all satisfy the marker)
//                        followed by fem::locate_dofs_topological (the dofs of the facets' closure),
//                        poisson_problem.cpp:56-75, elasticity_problem.cpp:124-142. One thread per
//                        (cell, local facet); every marked facet stores 1 for its closure dofs
//                        (Basix layout: vertices, edges, faces). Writers of one byte store the same value.
//   interpolate_source   the interpolation lambdas at the dof coordinates: Poisson
//                        f = 10 exp(-((x-.5)^2 + (y-.5)^2) / 0.02), g = sin(5x)
//                        (poisson_problem.cpp:84-105); elasticity f = (-dz r y, 1, dx r y),
//                        r = sqrt(dx^2 + dz^2) (elasticity_problem.cpp:154-175). Products and sums
//                        are spelled with the non-contracting intrinsics, so the arguments of
//                        exp / sin / sqrt are the host's bit for bit and the results differ from a
//                        host libm by the functions' own rounding only (<= 2 ulp).
// Run on the B200 since round 2 (GPU tests against the stand-in's bc_dofs / f / g); tests/emu also runs
// these sources on the host.
#include "kernels.h"

namespace ptb
{
namespace
{

constexpr int PD_THREADS = 256;

__device__ __forceinline__ bool dirichlet_marker(int problem, double x0, double x1)
{
  constexpr double eps = 1.0e-8;
  return problem == PTB_POISSON ? (fabs(x0) < eps || fabs(x0 - 1) < eps) : fabs(x1) < eps;
}

__global__ void locate_bc_facets(std::int64_t n_cells, int problem, int order, int nd,
                                 const double* __restrict__ xyz, const std::int32_t* __restrict__ x_dofmap,
                                 const std::int32_t* __restrict__ dofmap, std::uint8_t* __restrict__ bc)
{
  const std::int64_t k = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (k >= n_cells * 4)
    return;
  const std::int64_t cell = k >> 2;
  const int lf = static_cast<int>(k & 3);
  for (int v = 0; v < 4; ++v)
    if (v != lf)
    {
      const std::int64_t vert = x_dofmap[cell * 4 + v];
      if (!dirichlet_marker(problem, xyz[vert * 4], xyz[vert * 4 + 1]))
        return;
    }
  const std::int32_t* d = dofmap + cell * nd;
  const int ne = order - 1, nf = (order - 1) * (order - 2) / 2;
  for (int v = 0; v < 4; ++v)
    if (v != lf)
      bc[d[v]] = 1;
  // reference tetrahedron edges e = 0..5: (2,3) (1,3) (1,2) (0,3) (0,2) (0,1)
  constexpr int ea[6] = {2, 1, 1, 0, 0, 0}, eb[6] = {3, 3, 2, 3, 2, 1};
#pragma unroll
  for (int e = 0; e < 6; ++e)
    if (ea[e] != lf && eb[e] != lf)
      for (int s = 0; s < ne; ++s)
        bc[d[4 + e * ne + s]] = 1;
  for (int s = 0; s < nf; ++s)
    bc[d[4 + 6 * ne + lf * nf + s]] = 1;
}

// X: coordinates by dof, `stride` doubles apart (3 = caller's dof_x, 4 = the padded xdof of P1)
__global__ void interpolate_source(std::int64_t n, int problem, const double* __restrict__ X, int stride,
                                   double* __restrict__ f, double* __restrict__ g)
{
  const std::int64_t p = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (p >= n)
    return;
  const double x0 = X[p * stride], x1 = X[p * stride + 1], x2 = X[p * stride + 2];
  if (problem == PTB_POISSON)
  {
    const double dx = x0 - 0.5, dy = x1 - 0.5;
    const double dr = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    f[p] = 10 * exp(-dr / 0.02);
    g[p] = sin(5 * x0);
  }
  else
  {
    const double dx = x0 - 0.5, dz = x2 - 0.5;
    const double r = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dz, dz)));
    f[3 * p + 0] = __dmul_rn(__dmul_rn(-dz, r), x1);
    f[3 * p + 1] = 1.0;
    f[3 * p + 2] = __dmul_rn(__dmul_rn(dx, r), x1);
  }
}

} // namespace

#ifndef PTB_HOST_EMU // launchers: device build only
void launch_locate_bc(ptb_ctx* c)
{
  const std::int64_t n = c->n_cells * 4;
  c->bc.zero(c->stream);
  locate_bc_facets<<<static_cast<unsigned>((n + PD_THREADS - 1) / PD_THREADS), PD_THREADS, 0, c->stream>>>(
      c->n_cells, c->problem, c->order, c->nd, c->xyz.p, c->x_dofmap.p, c->dofmap.p, c->bc.p);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_interpolate_source(ptb_ctx* c, const double* X, int stride)
{
  const std::int64_t n = static_cast<std::int64_t>(c->n_owned) + c->n_ghost;
  interpolate_source<<<static_cast<unsigned>((n + PD_THREADS - 1) / PD_THREADS), PD_THREADS, 0, c->stream>>>(
      n, c->problem, X, stride, c->f.p, c->g.p);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}
#endif // PTB_HOST_EMU

} // namespace ptb
