// Host stand-in for the mesh layer the reference gets from DOLFINx.
//
// Mirrors /root/reference/src/mesh.cpp:
//   num_entities / num_pdofs      mesh.cpp:44-74
//   create_cube_mesh sizing       mesh.cpp:78-151   (replayed exactly, integer arithmetic)
//   mesh::create_box(tetrahedron) mesh.cpp:184-186  (Kuhn 6-tet split)
//
// What is deliberately different (DESIGN.md "Out of scope"):
//   * no graph partitioner: ranks own contiguous z-slabs of cube layers;
//   * r uniform refinements are replaced by generating the (N << r) box directly (same entity
//     counts, mesh.cpp:44-54) and the summary line says so;
//   * every rank carries one ghost layer of cells below its slab so that every owned dof sees all
//     of its cells locally (DOLFINx GhostMode::shared_vertex instead of ::none).
#pragma once
#include <array>
#include <cstdint>
#include <vector>

namespace ptb::host
{

/// (vertices, edges, faces, cells) of the i x j x k Kuhn box after nrefine dyadic refinements
/// (mesh.cpp:44-54).
std::array<std::int64_t, 4> num_entities(std::int64_t i, std::int64_t j, std::int64_t k,
                                         int nrefine);

/// Scalar Lagrange dofs of the given order (mesh.cpp:56-74). Throws for order not in 1..4.
std::int64_t num_pdofs(std::int64_t i, std::int64_t j, std::int64_t k, int nrefine, int order);

struct CubeSizing
{
  std::int64_t Nx, Ny, Nz;
  int r;
};

/// Replay of create_cube_mesh's choice of (Nx, Ny, Nz, r) (mesh.cpp:86-151).
CubeSizing cube_mesh_sizing(std::size_t target_dofs, bool target_dofs_total,
                            std::size_t dofs_per_node, int order, std::size_t num_processes);

/// Vertex offsets (bit 0 = x, bit 1 = y, bit 2 = z) of the six tetrahedra of a cube, in the cell
/// order and local vertex order of SURVEY Appendix B1.
extern const int kuhn_tets[6][4];

/// The local part of the nx x ny x nz unit-cube tetrahedral mesh held by one rank.
struct BoxMesh
{
  std::int64_t nx = 0, ny = 0, nz = 0; // global cube counts (already << r)
  int rank = 0, nranks = 1;
  // Cube layers [L0, L1) are owned; layers [l0, l1) are local (l0 = L0 - 1 on ranks > 0).
  std::int64_t L0 = 0, L1 = 0, l0 = 0, l1 = 0;
  // Vertex planes [P0, P1) are owned (P1 = nz + 1 on the last rank, else L1).
  std::int64_t P0 = 0, P1 = 0;

  std::int64_t n_cells_local() const { return 6 * nx * ny * (l1 - l0); }
  std::int64_t n_cells_owned() const { return 6 * nx * ny * (L1 - L0); }
  std::int64_t n_cells_global() const { return 6 * nx * ny * nz; }
  std::int64_t n_ghost_cells_front() const { return 6 * nx * ny * (L0 - l0); }
  std::int64_t n_vertices_local() const { return (nx + 1) * (ny + 1) * (l1 - l0 + 1); }
  std::int64_t cell_global_offset() const { return 6 * nx * ny * l0; }

  std::vector<double> x;               // [n_vertices_local * 3], planes l0..l1, iy, ix
  std::vector<std::int32_t> x_dofmap;  // [n_cells_local * 4]
};

/// Layers owned by `rank` when nz layers are split as evenly as possible over nranks slabs.
std::array<std::int64_t, 2> slab_range(std::int64_t nz, int rank, int nranks);

/// Build the local slab (geometry + cell->vertex map). with_arrays = false fills the sizes and ranges
/// only (x and x_dofmap stay empty): for callers that generate the arrays on the device.
BoxMesh create_box_mesh(std::int64_t nx, std::int64_t ny, std::int64_t nz, int rank, int nranks,
                        bool with_arrays = true);

} // namespace ptb::host
