/* C API of the host stand-in (libptb200_host.so).
 *
 * The reference (FEniCS/performance-test) builds its mesh, function space, boundary conditions,
 * source terms and sparsity pattern with DOLFINx on the host *before* the timed hot regions
 * (src/main.cpp:130-170, src/poisson_problem.cpp:33-123, src/elasticity_problem.cpp:101-197).
 * DOLFINx is not available here, so this library produces the same kind of arrays for the
 * reference's unit-cube tetrahedral mesh. Its outputs are the inputs of the CUDA C-ABI in
 * ptb200.h; the oracle (tests only) consumes the same arrays.
 *
 * All functions return 0 on success; on failure they return non-zero and pth_last_error() gives
 * the message (the reference throws std::runtime_error in the same situations).
 */
#ifndef PTB200_HOST_H
#define PTB200_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pth_problem pth_problem;

const char* pth_last_error(void);

/* src/mesh.cpp:44-54 and :56-74 */
int pth_num_entities(int64_t i, int64_t j, int64_t k, int nrefine, int64_t out4[4]);
int pth_num_pdofs(int64_t i, int64_t j, int64_t k, int nrefine, int order, int64_t* out);

/* create_cube_mesh's choice of (Nx, Ny, Nz, r), src/mesh.cpp:78-151. */
int pth_cube_sizing(uint64_t target_dofs, int target_dofs_total, uint64_t dofs_per_node, int order,
                    uint64_t num_processes, int64_t out4[4]);

/* Build mesh slab + function space + BC + source terms + sparsity pattern for one rank.
 * problem_type: "poisson" | "elasticity" (cgpoisson uses the poisson data). */
int pth_problem_create(const char* problem_type, int order, int64_t nx, int64_t ny, int64_t nz,
                       int rank, int nranks, pth_problem** out);
/* The same object with the surface-sized data only (sizes, ghost and halo lists, exterior facets):
 * for callers that generate mesh, dofmap, pattern, Dirichlet dofs and sources on the device
 * (ptb_create_box, ptb_build_pattern, ptb_locate_bc, ptb_interpolate_source). x, x_dofmap, dofmap,
 * dof_x, rowptr/cols, bc_dofs, f, g stay empty; nnz and n_bc read 0. */
int pth_problem_create_sizes_only(const char* problem_type, int order, int64_t nx, int64_t ny,
                                  int64_t nz, int rank, int nranks, pth_problem** out);
/* Renumber the owned dofs of a single-rank problem the way a real DOLFINx dofmap is numbered: not
 * lattice-lexicographically. kind = "rcm" (reverse Cuthill-McKee over the sparsity graph: banded,
 * local, not translation invariant -- what fem::DofMap's graph reordering produces in kind) or
 * "random" (seeded shuffle: no locality at all). dofmap, dof_x, f, g, bc_dofs, rowptr and cols follow. */
int pth_problem_renumber(pth_problem* p, const char* kind, uint64_t seed);
void pth_problem_destroy(pth_problem* p);

/* Named scalar: n_cells, n_cells_owned, n_ghost_cells_front, cell_global_offset, n_cells_global,
 * n_vertices, nd, bs, order, n_owned, n_ghost, n_global, global_offset, nnz, n_bc, n_facets,
 * n_nbr, rank, nranks, nx, ny, nz. */
int pth_problem_scalar(const pth_problem* p, const char* name, int64_t* out);

/* Named array view (owned by the problem). dtype: 0 f64, 1 i32, 2 i64. Names: x, x_dofmap, dofmap,
 * dof_x, ghost_global, ghost_owner, rowptr, cols, bc_dofs, f, g, facet_cells, facet_local,
 * nbr_ranks, send_displ, recv_displ, local_indices, remote_indices. */
int pth_problem_array(const pth_problem* p, const char* name, const void** data, int64_t* count,
                      int* dtype);

#ifdef __cplusplus
}
#endif
#endif
