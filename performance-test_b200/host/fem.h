// Host stand-in for the DOLFINx objects the reference's problem files build before the timed hot
// regions: FunctionSpace/DofMap (poisson_problem.cpp:33-47), DirichletBC dof location (:51-79),
// RHS interpolation (:82-108), sparsity pattern (:122-123) and the exterior-facet list the
// g*v*ds integral of Poisson.py:32 runs over. Elasticity analogues: elasticity_problem.cpp:101-179.
//
// Conventions (Basix/UFC):
//   reference tet edges  e0=(2,3) e1=(1,3) e2=(1,2) e3=(0,3) e4=(0,2) e5=(0,1)
//   reference tet facets f = face opposite local vertex f
//   Lagrange (gll_warped) local dofs: 4 vertices, then (k-1) per edge, then (k-1)(k-2)/2 per face
//   edge dofs are ordered along the edge from its lower *global* vertex (orientation baked into
//   the dofmap, no per-cell transformation)
#pragma once
#include "box_mesh.h"
#include <cstdint>
#include <string>
#include <vector>

namespace ptb::host
{

extern const int tet_edges[6][2];
extern const int tet_faces[4][3];

inline int lagrange_ndofs(int order) { return (order + 1) * (order + 2) * (order + 3) / 6; }

/// Degree-k scalar Lagrange space (block size bs = 1 Poisson, 3 elasticity) on the local slab.
struct FunctionSpace
{
  int order = 1, bs = 1, nd = 4;
  std::int64_t n_global = 0;      // global block dofs
  std::int64_t global_offset = 0; // global index of local dof 0
  std::int32_t n_owned = 0, n_ghost = 0, n_ghost_low = 0, n_ghost_high = 0;
  std::vector<std::int64_t> ghost_global; // [n_ghost]
  std::vector<std::int32_t> ghost_owner;  // [n_ghost]
  std::vector<std::int32_t> dofmap;       // [n_cells_local * nd], local block indices
  std::vector<double> dof_x;              // [(n_owned + n_ghost) * 3] dof coordinates
  // Halo (forward scatter owner -> ghost), DOLFINx Scatterer layout (cgpoisson_problem.cpp:187-229):
  // neighbour ranks, per-neighbour displacements, owned indices to send, ghost positions to fill.
  std::vector<std::int32_t> nbr_ranks;
  std::vector<std::int32_t> send_displ, recv_displ;         // [n_nbr + 1]
  std::vector<std::int32_t> local_indices, remote_indices;  // block indices
};

/// with_dofmap = false fills the sizes, ghost and halo lists only (dofmap and dof_x stay empty): for
/// callers that generate the dofmap on the device.
FunctionSpace create_functionspace(const BoxMesh& mesh, int order, int bs, bool with_dofmap = true);

/// Block dofs (owned and ghost, ascending) on the Dirichlet boundary:
/// "poisson": x = 0 or x = 1 (poisson_problem.cpp:60-71); "elasticity": y = 0
/// (elasticity_problem.cpp:127-138).
std::vector<std::int32_t> locate_bc_dofs(const BoxMesh& mesh, const FunctionSpace& V,
                                         const std::string& problem);

/// Nodal interpolation of the reference's source terms at the dof points.
/// poisson: f = 10 exp(-((x-.5)^2 + (y-.5)^2)/0.02), g = sin(5x)   (poisson_problem.cpp:86-106)
/// elasticity: f = (-dz r y, 1, dx r y), r = sqrt(dx^2+dz^2)       (elasticity_problem.cpp:155-176)
void interpolate_rhs(const FunctionSpace& V, const std::string& problem, std::vector<double>& f,
                     std::vector<double>& g);

/// (cell, local_facet) pairs of all boundary triangles of local cells, ascending in cell.
void exterior_facets(const BoxMesh& mesh, std::vector<std::int32_t>& cells,
                     std::vector<std::int32_t>& local_facets);

} // namespace ptb::host
