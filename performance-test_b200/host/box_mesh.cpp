#include "box_mesh.h"
#include <cstdlib>
#include <stdexcept>

namespace ptb::host
{

std::array<std::int64_t, 4> num_entities(std::int64_t i, std::int64_t j, std::int64_t k,
                                         int nrefine)
{
  i <<= nrefine;
  j <<= nrefine;
  k <<= nrefine;
  const std::int64_t cubes = i * j * k;
  const std::int64_t s = i * j + i * k + j * k;
  const std::int64_t t = i + j + k;
  return {(i + 1) * (j + 1) * (k + 1), 7 * cubes + 3 * s + t, 12 * cubes + 2 * s, 6 * cubes};
}

std::int64_t num_pdofs(std::int64_t i, std::int64_t j, std::int64_t k, int nrefine, int order)
{
  const auto [nv, ne, nf, nc] = num_entities(i, j, k, nrefine);
  // Lagrange: 1 dof per vertex, (p-1) per edge, (p-1)(p-2)/2 per face, (p-1)(p-2)(p-3)/6 per cell.
  if (order == 1)
    return nv;
  if (order == 2)
    return nv + ne;
  if (order == 3)
    return nv + 2 * ne + nf;
  if (order == 4)
    return nv + 3 * ne + 3 * nf + nc;
  throw std::runtime_error("Order not supported");
}

CubeSizing cube_mesh_sizing(std::size_t target_dofs, bool target_dofs_total,
                            std::size_t dofs_per_node, int order, std::size_t num_processes)
{
  // Target number of scalar "nodes" (mesh.cpp:86-90): unsigned integer division, then signed.
  const std::int64_t N = target_dofs_total ? target_dofs / dofs_per_node
                                           : target_dofs * num_processes / dofs_per_node;

  // Cubic initial guess; past Nx_max = 200 the reference adds refinement levels until it
  // overshoots and then shrinks the base box back under the target (mesh.cpp:98-126).
  constexpr std::int64_t Nx_max = 200;
  std::int64_t n = 1;
  int r = 0;
  for (std::int64_t have = 0; have < N; have = num_pdofs(n, n, n, r, order))
  {
    ++n;
    if (n > Nx_max)
    {
      do
      {
        ++r;
        have = num_pdofs(n, n, n, r, order);
      } while (have < N);
      while (have > N)
      {
        --n;
        have = num_pdofs(n, n, n, r, order);
      }
    }
  }

  // Neighbourhood search (mesh.cpp:134-151): i in [n-10, n+10), j and k in [i-5, i+5); the
  // first strict improvement over 1e6 wins; otherwise the cubic guess stands.
  CubeSizing best{n, n, n, r};
  std::size_t mindiff = 1000000;
  for (std::int64_t i = n - 10; i < n + 10; ++i)
    for (std::int64_t j = i - 5; j < i + 5; ++j)
      for (std::int64_t k = i - 5; k < i + 5; ++k)
      {
        const std::size_t diff = std::llabs(num_pdofs(i, j, k, r, order) - N);
        if (diff < mindiff)
        {
          mindiff = diff;
          best = {i, j, k, r};
        }
      }
  return best;
}

// Cube corner c has offset (c & 1, (c >> 1) & 1, (c >> 2) & 1). All six tets share the body
// diagonal 0-7 (Kuhn / Freudenthal split).
const int kuhn_tets[6][4]
    = {{0, 1, 3, 7}, {0, 1, 7, 5}, {0, 5, 7, 4}, {0, 3, 2, 7}, {0, 6, 4, 7}, {0, 2, 6, 7}};

std::array<std::int64_t, 2> slab_range(std::int64_t nz, int rank, int nranks)
{
  const std::int64_t base = nz / nranks, rem = nz % nranks;
  const std::int64_t lo = rank * base + (rank < rem ? rank : rem);
  return {lo, lo + base + (rank < rem ? 1 : 0)};
}

BoxMesh create_box_mesh(std::int64_t nx, std::int64_t ny, std::int64_t nz, int rank, int nranks,
                        bool with_arrays)
{
  if (nx < 1 || ny < 1 || nz < 1)
    throw std::runtime_error("create_box_mesh: box dimensions must be positive");
  if (nranks < 1 || rank < 0 || rank >= nranks || nz < nranks)
    throw std::runtime_error("create_box_mesh: need 0 <= rank < nranks <= nz");
  BoxMesh m;
  m.nx = nx, m.ny = ny, m.nz = nz, m.rank = rank, m.nranks = nranks;
  const auto [L0, L1] = slab_range(nz, rank, nranks);
  m.L0 = L0, m.L1 = L1;
  m.l0 = rank > 0 ? L0 - 1 : L0;
  m.l1 = L1;
  m.P0 = L0;
  m.P1 = rank == nranks - 1 ? nz + 1 : L1;

  const std::int64_t nvx = nx + 1, nvy = ny + 1, nvp = nvx * nvy;
  if (m.n_vertices_local() > INT32_MAX || m.n_cells_local() * 4 > (std::int64_t)UINT32_MAX)
    throw std::runtime_error("create_box_mesh: local slab exceeds 32-bit local indexing");
  if (!with_arrays)
    return m;

  // Geometry as DOLFINx's create_box lays it out: x = a + ix * ((b - a) / nx), a = 0, b = 1.
  const double hx = 1.0 / static_cast<double>(nx), hy = 1.0 / static_cast<double>(ny),
               hz = 1.0 / static_cast<double>(nz);
  m.x.resize(static_cast<std::size_t>(m.n_vertices_local()) * 3);
#pragma omp parallel for schedule(static)
  for (std::int64_t pz = m.l0; pz <= m.l1; ++pz)
    for (std::int64_t iy = 0; iy < nvy; ++iy)
      for (std::int64_t ix = 0; ix < nvx; ++ix)
      {
        double* p = m.x.data() + 3 * ((pz - m.l0) * nvp + iy * nvx + ix);
        p[0] = 0.0 + hx * static_cast<double>(ix);
        p[1] = 0.0 + hy * static_cast<double>(iy);
        p[2] = 0.0 + hz * static_cast<double>(pz);
      }

  m.x_dofmap.resize(static_cast<std::size_t>(m.n_cells_local()) * 4);
#pragma omp parallel for schedule(static)
  for (std::int64_t iz = m.l0; iz < m.l1; ++iz)
    for (std::int64_t iy = 0; iy < ny; ++iy)
      for (std::int64_t ix = 0; ix < nx; ++ix)
      {
        const std::int64_t v0 = (iz - m.l0) * nvp + iy * nvx + ix;
        std::int32_t* c = m.x_dofmap.data() + 24 * (((iz - m.l0) * ny + iy) * nx + ix);
        for (int t = 0; t < 6; ++t)
          for (int a = 0; a < 4; ++a)
          {
            const int o = kuhn_tets[t][a];
            c[4 * t + a] = static_cast<std::int32_t>(v0 + (o & 1) + ((o >> 1) & 1) * nvx
                                                     + ((o >> 2) & 1) * nvp);
          }
      }
  return m;
}

} // namespace ptb::host
