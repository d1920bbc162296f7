// The structured entity numbering of the Kuhn box and the per-tetrahedron-type local dof table of
// the Lagrange P1-P3 spaces on it: one definition for the host stand-in (host/fem.cpp) and for the
// device-side dofmap generator (csrc/box.cu). Conventions: host/fem.h.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace ptb::kuhn
{

// Cube corner c has offset (c & 1, (c >> 1) & 1, (c >> 2) & 1). All six tets share the body
// diagonal 0-7 (Kuhn / Freudenthal split).
inline constexpr int kuhn_tets[6][4]
    = {{0, 1, 3, 7}, {0, 1, 7, 5}, {0, 5, 7, 4}, {0, 3, 2, 7}, {0, 6, 4, 7}, {0, 2, 6, 7}};
inline constexpr int tet_edges[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
inline constexpr int tet_faces[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};

inline int lagrange_ndofs(int order) { return (order + 1) * (order + 2) * (order + 3) / 6; }

// Structured entity "kinds" of the Kuhn box. An entity is (kind, base lattice point); the other
// vertices sit at base + offset code (bit 0 = x, 1 = y, 2 = z). Kinds whose offsets have no z
// bit live in a vertex plane, the others in the cube layer above the base plane.
struct Kind
{
  int dim;    // 0 vertex, 1 edge, 2 face
  int d1, d2; // offset codes (d2 = 0 unless face)
  bool layer; // false: plane block, true: layer block
  int ex, ey; // lattice extent in x, y (bases run over (nx+1-ex) x (ny+1-ey))
};

constexpr int NK = 20;
inline constexpr Kind kinds[NK] = {
    {0, 0, 0, false, 0, 0},                                                   // 0 vertices
    {1, 1, 0, false, 1, 0}, {1, 2, 0, false, 0, 1}, {1, 3, 0, false, 1, 1},   // 1-3 in-plane edges
    {2, 1, 3, false, 1, 1}, {2, 2, 3, false, 1, 1},                           // 4-5 in-plane faces
    {1, 4, 0, true, 0, 0},  {1, 5, 0, true, 1, 0},                            // 6-9 rising edges
    {1, 6, 0, true, 0, 1},  {1, 7, 0, true, 1, 1},
    {2, 3, 7, true, 1, 1},  {2, 1, 7, true, 1, 1},  {2, 5, 7, true, 1, 1},    // 10-19 rising faces
    {2, 4, 7, true, 1, 1},  {2, 6, 7, true, 1, 1},  {2, 2, 7, true, 1, 1},
    {2, 2, 6, true, 0, 1},  {2, 4, 6, true, 0, 1},  {2, 1, 5, true, 1, 0},
    {2, 4, 5, true, 1, 0}};

inline int find_kind(int dim, int d1, int d2)
{
  for (int k = 0; k < NK; ++k)
    if (kinds[k].dim == dim && kinds[k].d1 == d1 && kinds[k].d2 == d2)
      return k;
  throw std::runtime_error("fem: unknown entity kind in Kuhn split");
}

inline int dofs_per_entity(int dim, int order)
{
  if (dim == 0)
    return 1;
  if (dim == 1)
    return order - 1;
  return (order - 1) * (order - 2) / 2;
}

// Layout of the global numbering: level-major; each level is [plane block][layer block]; inside a
// block kinds are contiguous, lexicographic (iy, ix), entity sub-dofs adjacent.
struct Numbering
{
  std::int64_t nx, ny, nz;
  int order;
  std::int64_t koff[NK]; // offset of kind inside its block
  std::int64_t kw[NK];   // bases per lattice row
  int ksub[NK];
  std::int64_t PS = 0, LS = 0;

  Numbering(std::int64_t nx_, std::int64_t ny_, std::int64_t nz_, int order_)
      : nx(nx_), ny(ny_), nz(nz_), order(order_)
  {
    for (int k = 0; k < NK; ++k)
    {
      const Kind& K = kinds[k];
      ksub[k] = dofs_per_entity(K.dim, order);
      kw[k] = nx + 1 - K.ex;
      const std::int64_t count = kw[k] * (ny + 1 - K.ey) * ksub[k];
      std::int64_t& S = K.layer ? LS : PS;
      koff[k] = S;
      S += count;
    }
  }
  std::int64_t level_stride() const { return PS + LS; }
  std::int64_t total() const { return nz * (PS + LS) + PS; }
  std::int64_t global(int k, std::int64_t level, std::int64_t iy, std::int64_t ix, int sub) const
  {
    return level * (PS + LS) + (kinds[k].layer ? PS : 0) + koff[k] + (iy * kw[k] + ix) * ksub[k]
           + sub;
  }
};

// Per (tet type, local dof): which entity it sits on, relative to the cube's corner 0.
struct LocalDof
{
  int kind, bx, by, bz, sub;
};

inline void build_local_table(int order, std::vector<LocalDof>& tab)
{
  const int nd = lagrange_ndofs(order);
  const int ne = order - 1, nf = (order - 1) * (order - 2) / 2;
  tab.assign(6 * nd, LocalDof{});
  for (int t = 0; t < 6; ++t)
  {
    LocalDof* T = tab.data() + t * nd;
    for (int a = 0; a < 4; ++a)
    {
      const int c = kuhn_tets[t][a];
      T[a] = {0, c & 1, (c >> 1) & 1, (c >> 2) & 1, 0};
    }
    for (int e = 0; e < 6; ++e)
    {
      const int ca = kuhn_tets[t][tet_edges[e][0]], cb = kuhn_tets[t][tet_edges[e][1]];
      int base, tip;
      bool agree; // local low->high vertex order equals global base->tip
      if ((ca & cb) == ca)
        base = ca, tip = cb, agree = true;
      else if ((ca & cb) == cb)
        base = cb, tip = ca, agree = false;
      else
        throw std::runtime_error("fem: Kuhn edge is not monotone");
      const int k = find_kind(1, tip ^ base, 0);
      for (int s = 0; s < ne; ++s)
        T[4 + e * ne + s]
            = {k, base & 1, (base >> 1) & 1, (base >> 2) & 1, agree ? s : ne - 1 - s};
    }
    for (int f = 0; f < 4 && nf > 0; ++f)
    {
      const int c0 = kuhn_tets[t][tet_faces[f][0]], c1 = kuhn_tets[t][tet_faces[f][1]],
                c2 = kuhn_tets[t][tet_faces[f][2]];
      const int base = c0 & c1 & c2;
      if (base != c0 && base != c1 && base != c2)
        throw std::runtime_error("fem: Kuhn face has no minimal corner");
      int d[2], n = 0;
      for (int c : {c0, c1, c2})
        if (c != base)
          d[n++] = c ^ base;
      if (d[0] > d[1])
        std::swap(d[0], d[1]);
      const int k = find_kind(2, d[0], d[1]);
      for (int s = 0; s < nf; ++s)
        T[4 + 6 * ne + f * nf + s] = {k, base & 1, (base >> 1) & 1, (base >> 2) & 1, s};
    }
  }
}

// GLL-warped edge parameters (distance from the edge's lower global vertex).
inline double edge_param(int order, int sub)
{
  if (order == 2)
    return 0.5;
  const double a = 0.5 * (1.0 - 1.0 / std::sqrt(5.0));
  return sub == 0 ? a : 1.0 - a;
}

} // namespace ptb::kuhn
