#include "fem.h"
#include "../common/kuhn_space.h"
#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace ptb::host
{

const int tet_edges[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
const int tet_faces[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
static_assert(ptb::kuhn::tet_edges[0][0] == 2 && ptb::kuhn::tet_faces[3][2] == 2, "conventions of common/kuhn_space.h");

namespace
{
using namespace ptb::kuhn; // Kind, kinds, Numbering, LocalDof, build_local_table, edge_param

// Local numbering: owned range [G0, G1) first, then the low ghost range [Glow, G0), then the high
// ghost range [G1, Ghigh).
struct LocalRanges
{
  std::int64_t G0, G1, Glow, Ghigh;
  std::int32_t to_local(std::int64_t g) const
  {
    if (g >= G0 && g < G1)
      return static_cast<std::int32_t>(g - G0);
    if (g >= Glow && g < G0)
      return static_cast<std::int32_t>((G1 - G0) + (g - Glow));
    if (g >= G1 && g < Ghigh)
      return static_cast<std::int32_t>((G1 - G0) + (G0 - Glow) + (g - G1));
    return -1;
  }
};

LocalRanges local_ranges(const BoxMesh& m, const Numbering& N)
{
  LocalRanges R;
  const std::int64_t S = N.level_stride();
  R.G0 = m.P0 * S;
  R.G1 = (m.rank == m.nranks - 1) ? N.total() : m.L1 * S;
  R.Glow = m.l0 * S; // == G0 on rank 0
  R.Ghigh = (m.rank == m.nranks - 1) ? R.G1 : R.G1 + N.PS;
  return R;
}
} // namespace

FunctionSpace create_functionspace(const BoxMesh& m, int order, int bs, bool with_dofmap)
{
  if (order < 1 || order > 3)
    throw std::runtime_error("Order not supported");
  FunctionSpace V;
  V.order = order, V.bs = bs, V.nd = lagrange_ndofs(order);
  const Numbering N(m.nx, m.ny, m.nz, order);
  const LocalRanges R = local_ranges(m, N);
  V.n_global = N.total();
  V.global_offset = R.G0;
  if (R.Ghigh - R.Glow > INT32_MAX)
    throw std::runtime_error("create_functionspace: local dof count exceeds int32");
  V.n_owned = static_cast<std::int32_t>(R.G1 - R.G0);
  V.n_ghost_low = static_cast<std::int32_t>(R.G0 - R.Glow);
  V.n_ghost_high = static_cast<std::int32_t>(R.Ghigh - R.G1);
  V.n_ghost = V.n_ghost_low + V.n_ghost_high;

  V.ghost_global.resize(V.n_ghost);
  V.ghost_owner.resize(V.n_ghost);
  for (std::int32_t i = 0; i < V.n_ghost_low; ++i)
    V.ghost_global[i] = R.Glow + i, V.ghost_owner[i] = m.rank - 1;
  for (std::int32_t i = 0; i < V.n_ghost_high; ++i)
    V.ghost_global[V.n_ghost_low + i] = R.G1 + i, V.ghost_owner[V.n_ghost_low + i] = m.rank + 1;

  // Halo lists. The lower neighbour owns [Glow, G0) as the tail of its range and wants our first
  // plane block; the upper neighbour owns [G1, Ghigh) as the head of its range and wants our last
  // (plane + layer) level.
  V.send_displ = {0};
  V.recv_displ = {0};
  const std::int32_t S = static_cast<std::int32_t>(N.level_stride());
  const std::int32_t PS = static_cast<std::int32_t>(N.PS);
  if (m.rank > 0)
  {
    V.nbr_ranks.push_back(m.rank - 1);
    for (std::int32_t i = 0; i < PS; ++i)
      V.local_indices.push_back(i);
    for (std::int32_t i = 0; i < V.n_ghost_low; ++i)
      V.remote_indices.push_back(V.n_owned + i);
    V.send_displ.push_back(static_cast<std::int32_t>(V.local_indices.size()));
    V.recv_displ.push_back(static_cast<std::int32_t>(V.remote_indices.size()));
  }
  if (m.rank < m.nranks - 1)
  {
    V.nbr_ranks.push_back(m.rank + 1);
    for (std::int32_t i = V.n_owned - S; i < V.n_owned; ++i)
      V.local_indices.push_back(i);
    for (std::int32_t i = 0; i < V.n_ghost_high; ++i)
      V.remote_indices.push_back(V.n_owned + V.n_ghost_low + i);
    V.send_displ.push_back(static_cast<std::int32_t>(V.local_indices.size()));
    V.recv_displ.push_back(static_cast<std::int32_t>(V.remote_indices.size()));
  }

  if (!with_dofmap)
    return V;

  // Cell dofmap from the per-tet-type table.
  std::vector<LocalDof> tab;
  build_local_table(order, tab);
  const int nd = V.nd;
  V.dofmap.resize(static_cast<std::size_t>(m.n_cells_local()) * nd);
  bool bad = false;
#pragma omp parallel for schedule(static) reduction(|| : bad)
  for (std::int64_t iz = m.l0; iz < m.l1; ++iz)
    for (std::int64_t iy = 0; iy < m.ny; ++iy)
      for (std::int64_t ix = 0; ix < m.nx; ++ix)
      {
        std::int32_t* cd
            = V.dofmap.data() + 6 * nd * (((iz - m.l0) * m.ny + iy) * m.nx + ix);
        for (int t = 0; t < 6; ++t)
          for (int i = 0; i < nd; ++i)
          {
            const LocalDof& L = tab[t * nd + i];
            const std::int64_t g = N.global(L.kind, iz + L.bz, iy + L.by, ix + L.bx, L.sub);
            const std::int32_t l = R.to_local(g);
            bad = bad || l < 0;
            cd[t * nd + i] = l;
          }
      }
  if (bad)
    throw std::runtime_error("create_functionspace: cell dof outside the local ranges");

  // Dof coordinates, by walking every local (level, kind, base, sub).
  V.dof_x.assign(static_cast<std::size_t>(V.n_owned + V.n_ghost) * 3, 0.0);
  const double hx = 1.0 / static_cast<double>(m.nx), hy = 1.0 / static_cast<double>(m.ny),
               hz = 1.0 / static_cast<double>(m.nz);
  for (std::int64_t level = m.l0; level <= m.l1; ++level)
    for (int k = 0; k < NK; ++k)
    {
      const Kind& K = kinds[k];
      if (N.ksub[k] == 0 || (K.layer && level == m.l1))
        continue;
      const std::int64_t W = N.kw[k], H = m.ny + 1 - K.ey;
#pragma omp parallel for schedule(static)
      for (std::int64_t iy = 0; iy < H; ++iy)
        for (std::int64_t ix = 0; ix < W; ++ix)
          for (int s = 0; s < N.ksub[k]; ++s)
          {
            const std::int32_t l = R.to_local(N.global(k, level, iy, ix, s));
            if (l < 0)
              continue; // level block not held by this rank (e.g. layer l1 - 1 above plane l1)
            double ox = 0, oy = 0, oz = 0; // offset from the base in lattice units
            if (K.dim == 1)
            {
              const double t = edge_param(order, s);
              ox = t * (K.d1 & 1), oy = t * ((K.d1 >> 1) & 1), oz = t * ((K.d1 >> 2) & 1);
            }
            else if (K.dim == 2)
            {
              ox = ((K.d1 & 1) + (K.d2 & 1)) / 3.0;
              oy = (((K.d1 >> 1) & 1) + ((K.d2 >> 1) & 1)) / 3.0;
              oz = (((K.d1 >> 2) & 1) + ((K.d2 >> 2) & 1)) / 3.0;
            }
            double* p = V.dof_x.data() + 3 * static_cast<std::int64_t>(l);
            p[0] = hx * (static_cast<double>(ix) + ox);
            p[1] = hy * (static_cast<double>(iy) + oy);
            p[2] = hz * (static_cast<double>(level) + oz);
          }
    }
  return V;
}

std::vector<std::int32_t> locate_bc_dofs(const BoxMesh& m, const FunctionSpace& V,
                                         const std::string& problem)
{
  const bool poisson = problem == "poisson" || problem == "cgpoisson";
  if (!poisson && problem != "elasticity")
    throw std::runtime_error("Unknown problem type: " + problem);
  const Numbering N(m.nx, m.ny, m.nz, V.order);
  const LocalRanges R = local_ranges(m, N);
  std::vector<std::int32_t> out;
  for (std::int64_t level = m.l0; level <= m.l1; ++level)
    for (int k = 0; k < NK; ++k)
    {
      const Kind& K = kinds[k];
      if (N.ksub[k] == 0 || (K.layer && level == m.l1))
        continue;
      // An entity lies in the plane x = const (y = const) iff its x (y) extent is zero; the
      // reference marks facets whose vertices all satisfy the predicate and takes their closure.
      if (poisson ? K.ex != 0 : K.ey != 0)
        continue;
      const std::int64_t W = N.kw[k], H = m.ny + 1 - K.ey;
      for (std::int64_t iy = 0; iy < H; ++iy)
        for (std::int64_t ix = 0; ix < W; ++ix)
        {
          const bool on = poisson ? (ix == 0 || ix == m.nx) : (iy == 0);
          if (!on)
            continue;
          for (int s = 0; s < N.ksub[k]; ++s)
          {
            const std::int32_t l = R.to_local(N.global(k, level, iy, ix, s));
            if (l >= 0)
              out.push_back(l);
          }
        }
    }
  std::sort(out.begin(), out.end());
  return out;
}

void interpolate_rhs(const FunctionSpace& V, const std::string& problem, std::vector<double>& f,
                     std::vector<double>& g)
{
  const std::int64_t n = static_cast<std::int64_t>(V.n_owned) + V.n_ghost;
  const double* X = V.dof_x.data();
  if (problem == "elasticity")
  {
    f.resize(static_cast<std::size_t>(n) * 3);
    g.clear();
#pragma omp parallel for schedule(static)
    for (std::int64_t p = 0; p < n; ++p)
    {
      const double dx = X[3 * p] - 0.5, dz = X[3 * p + 2] - 0.5;
      const double r = std::sqrt(dx * dx + dz * dz);
      f[3 * p + 0] = -dz * r * X[3 * p + 1];
      f[3 * p + 1] = 1.0;
      f[3 * p + 2] = dx * r * X[3 * p + 1];
    }
    return;
  }
  f.resize(static_cast<std::size_t>(n));
  g.resize(static_cast<std::size_t>(n));
#pragma omp parallel for schedule(static)
  for (std::int64_t p = 0; p < n; ++p)
  {
    const double dx = X[3 * p] - 0.5, dy = X[3 * p + 1] - 0.5;
    const double dr = dx * dx + dy * dy;
    f[p] = 10 * std::exp(-dr / 0.02);
    g[p] = std::sin(5 * X[3 * p]);
  }
}

void exterior_facets(const BoxMesh& m, std::vector<std::int32_t>& cells,
                     std::vector<std::int32_t>& local_facets)
{
  // For each (tet type, local facet): the cube face it lies in, as (axis, side) or axis = -1.
  int axis[6][4], side[6][4];
  for (int t = 0; t < 6; ++t)
    for (int f = 0; f < 4; ++f)
    {
      axis[t][f] = -1, side[t][f] = 0;
      for (int ax = 0; ax < 3; ++ax)
      {
        int sum = 0;
        for (int v = 0; v < 3; ++v)
          sum += (kuhn_tets[t][tet_faces[f][v]] >> ax) & 1;
        if (sum == 0 || sum == 3)
          axis[t][f] = ax, side[t][f] = sum / 3;
      }
    }
  cells.clear();
  local_facets.clear();
  const std::int64_t n[3] = {m.nx, m.ny, m.nz};
  for (std::int64_t iz = m.l0; iz < m.l1; ++iz)
    for (std::int64_t iy = 0; iy < m.ny; ++iy)
      for (std::int64_t ix = 0; ix < m.nx; ++ix)
      {
        const std::int64_t idx[3] = {ix, iy, iz};
        const bool touches = ix == 0 || ix == m.nx - 1 || iy == 0 || iy == m.ny - 1 || iz == 0
                             || iz == m.nz - 1;
        if (!touches)
          continue;
        const std::int64_t cube = ((iz - m.l0) * m.ny + iy) * m.nx + ix;
        for (int t = 0; t < 6; ++t)
          for (int f = 0; f < 4; ++f)
          {
            const int ax = axis[t][f];
            if (ax < 0)
              continue;
            if (idx[ax] == (side[t][f] ? n[ax] - 1 : 0))
            {
              cells.push_back(static_cast<std::int32_t>(6 * cube + t));
              local_facets.push_back(f);
            }
          }
      }
}

} // namespace ptb::host
