/* ptb200.h -- C ABI of libptb200.so: the B200 (sm_100a) hot path of FEniCS/performance-test.
 *
 * Drop-in boundary. The reference has no FFI of its own; its hot path sits in three
 * timed C++ regions that call DOLFINx/PETSc:
 *
 *   ZZZ Assemble matrix   src/poisson_problem.cpp:125-139   src/elasticity_problem.cpp:199-213
 *   ZZZ Assemble vector   src/poisson_problem.cpp:146-157   src/elasticity_problem.cpp:220-231
 *   ZZZ Solve             src/main.cpp:208-211 -> solver_function (src/poisson_problem.cpp:164-179,
 *                         src/elasticity_problem.cpp:246-261), with the CG loop of src/cg.h:38-86
 *
 * This header is what a maintainer binds instead of those bodies (INTEGRATION.md shows the C++
 * shim). Everything is plain pointers and sizes. Host arrays are borrowed for the duration of a
 * call and copied to the device; the context owns all device memory; there are no callbacks.
 * Every function returns 0 on success and a non-zero code on failure, in which case
 * ptb_last_error() returns the message (the C++ shim rethrows it as std::runtime_error, keeping
 * the reference's "uncaught exception => non-zero exit" convention, src/main.cpp:43,115,170).
 * Hot calls return only after their stream work (and NCCL work) has completed, so a host timer
 * stopped right after the call is truthful.
 *
 * There is no CPU fallback: without a CUDA device every device call fails with an error.
 * One context per GPU / per rank; calls on one context must come from one thread at a time.
 *
 * Index conventions are DOLFINx's: local block dof indices are int32, owned dofs first then
 * ghosts; vectors are [owned*bs | ghost*bs]; CSR row pointers are int64, columns int32 (local,
 * sorted ascending); blocked (bs = 3) matrices store 3x3 row-major blocks.
 */
#ifndef PTB200_H
#define PTB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ptb_ctx ptb_ctx;

enum { PTB_POISSON = 0, PTB_ELASTICITY = 1 };     /* src/Poisson.py, src/Elasticity.py */
enum { PTB_PC_NONE = 0, PTB_PC_JACOBI = 1 };      /* cg.h as is | z = D^-1 r */
enum { PTB_OP_ASSEMBLED = 0, PTB_OP_MATRIX_FREE = 1 };
enum { PTB_STAGE_ASSEMBLE_MATRIX = 0, PTB_STAGE_ASSEMBLE_VECTOR = 1, PTB_STAGE_SOLVE = 2,
       PTB_STAGE_SPMV = 3, PTB_STAGE_COUNT = 4 };

/* ---- lifecycle ------------------------------------------------------------------------- */
int ptb_create(int device, ptb_ctx** out);
void ptb_destroy(ptb_ctx* ctx);
/* Message of the last failed call on ctx (ctx may be NULL for ptb_create failures). */
const char* ptb_last_error(const ptb_ctx* ctx);
/* Run all work on an existing cudaStream_t (e.g. the caller's current stream) instead of the
 * context's own stream. */
int ptb_set_stream(ptb_ctx* ctx, void* cuda_stream);

/* ---- setup: what DOLFINx holds before the timed regions ----------------------------------- */
/* mesh->geometry().x() and .dofmap() (affine tets): x[n_vertices*3], x_dofmap[n_cells*4].
 * Cells must include every cell adjacent to an owned dof (DOLFINx ghost mode shared_vertex). */
int ptb_set_mesh(ptb_ctx* ctx, int64_t n_vertices, const double* x, int64_t n_cells,
                 const int32_t* x_dofmap);
/* V->dofmap(): Lagrange order 1..3 on tets (basix gll_warped, poisson_problem.cpp:35-38), block
 * size bs (1 Poisson, 3 elasticity), dofmap[n_cells * nd] local block indices. */
int ptb_set_space(ptb_ctx* ctx, int problem, int order, int bs, int32_t n_owned, int32_t n_ghost,
                  const int32_t* dofmap);
/* "ZZZ Create Mesh" + "ZZZ FunctionSpace" on the device for the reference's unit cube
 * (mesh.cpp:184-186 mesh::create_box(tetrahedron); poisson_problem.cpp:33-47,
 * elasticity_problem.cpp:100-112): replaces ptb_set_mesh + ptb_set_space. nx x ny x nz cubes, six
 * tetrahedra each; rank owns a z-slab of cube layers plus one ghost layer of cells below it (the
 * shared_vertex ghost mode the row-owner assembly needs); Lagrange order 1..3, dofs numbered
 * level-major by entity kind: owned levels first, then the ghost level below, then the ghost plane
 * above. sizes receives n_vertices, n_cells, n_owned, n_ghost. ptb_get_mesh / ptb_get_dofmap /
 * ptb_get_dof_coordinates copy the arrays back (x [n_vertices*3], x_dofmap [n_cells*4], dofmap
 * [n_cells*nd], dof_x [(n_owned+n_ghost)*3]; any may be NULL). */
int ptb_create_box(ptb_ctx* ctx, int problem, int bs, int order, int64_t nx, int64_t ny, int64_t nz,
                   int rank, int nranks, int64_t sizes[4]);
int ptb_get_mesh(ptb_ctx* ctx, double* x, int32_t* x_dofmap);
int ptb_get_dofmap(ptb_ctx* ctx, int32_t* dofmap);
int ptb_get_dof_coordinates(ptb_ctx* ctx, double* dof_x);
/* Sparsity pattern of the owned rows (fem::create_sparsity_pattern / create_matrix,
 * poisson_problem.cpp:122-123). Builds the cell -> CSR-slot map and the device layout. */
int ptb_set_pattern(ptb_ctx* ctx, const int64_t* rowptr, const int32_t* cols);
/* The same pattern built on the device from the dofmap of ptb_set_space instead of being passed in
 * (fem::create_sparsity_pattern + create_matrix, poisson_problem.cpp:122-123, which the reference
 * times inside "ZZZ Assemble matrix"): per owned row the ascending union of the dofs of its cells.
 * Equivalent to ptb_set_pattern with that pattern; *nnz receives rowptr[n_owned]. The caller reads
 * the CSR arrays back with ptb_get_pattern to create its own Mat (rowptr [n_owned + 1], cols [nnz];
 * either may be NULL). */
int ptb_build_pattern(ptb_ctx* ctx, int64_t* nnz);
int ptb_get_pattern(ptb_ctx* ctx, int64_t* rowptr, int32_t* cols);
/* bc->dof_indices(): constrained block dofs (owned and ghost); all bs components constrained,
 * boundary value 0 (u0 = 0, poisson_problem.cpp:53-54). */
int ptb_set_bc(ptb_ctx* ctx, int32_t n_bc, const int32_t* bc_dofs);
/* "ZZZ Create boundary conditions" on the device (poisson_problem.cpp:51-79,
 * elasticity_problem.cpp:117-146): marks the facets whose three vertices satisfy the reference's
 * predicate (Poisson |x0| < 1e-8 or |x0 - 1| < 1e-8; elasticity |x1| < 1e-8), constrains the dofs of
 * their closure, and replaces ptb_set_bc. *n_bc receives the number of constrained local block dofs;
 * ptb_get_bc copies them out in ascending order (bdofs of fem::locate_dofs_topological). */
int ptb_locate_bc(ptb_ctx* ctx, int32_t* n_bc);
int ptb_get_bc(ptb_ctx* ctx, int32_t* bc_dofs);
/* Exterior facets as (cell, local_facet) pairs for the g*v*ds term (Poisson.py:32). */
int ptb_set_exterior_facets(ptb_ctx* ctx, int64_t n_facets, const int32_t* cells,
                            const int32_t* local_facets);
/* Coefficients of L: f->x()->array() [(n_owned+n_ghost)*bs] and g (Poisson only, else NULL). */
int ptb_set_source(ptb_ctx* ctx, const double* f, const double* g);
/* "ZZZ Create RHS function" on the device (poisson_problem.cpp:82-108, elasticity_problem.cpp:150-178):
 * evaluates the reference's interpolation lambdas at the dof coordinates and replaces ptb_set_source.
 * dof_x [(n_owned+n_ghost)*3] = V->tabulate_dof_coordinates(); NULL for order 1, where the dofs sit
 * on the vertices the context already holds, and for a space generated by ptb_create_box. ptb_get_source copies f [(n_owned+n_ghost)*bs] and g
 * [(n_owned+n_ghost)] (Poisson; may be NULL) back. */
int ptb_interpolate_source(ptb_ctx* ctx, const double* dof_x);
int ptb_get_source(ptb_ctx* ctx, double* f, double* g);
/* Re-upload geometry coordinates only (same topology). */
int ptb_update_geometry(ptb_ctx* ctx, const double* x);
/* common::Scatterer lists (cgpoisson_problem.cpp:187-229): forward scatter owner -> ghost. */
int ptb_set_halo(ptb_ctx* ctx, int n_nbr, const int32_t* nbr_ranks, const int32_t* send_displ,
                 const int32_t* local_indices, const int32_t* recv_displ,
                 const int32_t* remote_indices);
/* NCCL bootstrap: rank 0 calls ptb_nccl_unique_id, the host broadcasts the 128 bytes (MPI_Bcast
 * in the reference's world, torch.distributed here), every rank calls ptb_comm_init. */
int ptb_nccl_unique_id(void* out128);
int ptb_comm_init(ptb_ctx* ctx, int rank, int nranks, const void* unique_id128);

/* NCCL-free alternative over NVLink peer memory (CUDA IPC): the CG kernels pull ghost values
 * straight out of the neighbours' vectors and all-reduce their dot products through per-rank
 * windows -- no collective launches inside the solve. Call after ptb_set_space + ptb_set_halo:
 * every rank exports 192 bytes (3 IPC handles), the host all-gathers them, every rank connects.
 * src_index[j] = index, in the OWNER's local numbering, of the dof received as remote_indices[j]
 * (i.e. the owner's local_indices entry that feeds it). Re-connect if the space is set again. */
int ptb_peer_export(ptb_ctx* ctx, void* handles192);
int ptb_peer_connect(ptb_ctx* ctx, int rank, int nranks, const void* all_handles,
                     const int32_t* src_index);

/* ---- hot calls -------------------------------------------------------------------------- */
/* ZZZ Assemble matrix: element kernels + BC row/col zeroing + unit BC diagonal, deterministic
 * (no atomics). Also extracts the Jacobi diagonal. */
int ptb_assemble_matrix(ptb_ctx* ctx);
/* ZZZ Assemble vector: cells (+ exterior facets), b[bc] = 0. Owned rows gather from all their
 * cells, so no reverse scatter is needed. */
int ptb_assemble_vector(ptb_ctx* ctx);
/* ZZZ Solve: linalg::cg (cg.h:38-86) with optional Jacobi; stop when |r|^2/|r0|^2 < rtol^2
 * (cg.h:78), at most kmax iterations. x starts from the initial guess (zero unless
 * ptb_set_initial_guess was called) and b is the assembled RHS (or ptb_set_rhs). */
int ptb_cg_solve(ptb_ctx* ctx, int kmax, double rtol, int precond, int* iterations,
                 double* rel_residual);
/* How ptb_cg_solve runs the loop of cg.h:57-84: 1 = one persistent cooperative kernel (grid barriers
 * instead of three launches per iteration; pays when the per-GPU problem is small, i.e. under
 * strong scaling), 0 = three kernels per iteration, -1 = automatic (default: a single GPU decides by
 * its row count; across GPUs the loop is only taken when the caller asks for it, because EVERY rank
 * must pass the same value -- decide from the global size, e.g. global DOFs / ranks <= 2 M).
 * Results are identical either way (same kernels' arithmetic, same reduction order). */
int ptb_set_cg_persistent(ptb_ctx* ctx, int mode);
/* Operator used by ptb_cg_solve / ptb_apply_operator: the assembled matrix (default) or, for
 * the scalar Poisson spaces (P1-P3), the matrix-free action of the reference's cgpoisson problem
 * (src/cgpoisson_problem.cpp:193-230, form M of src/Poisson.py:33): y = sum_cells Ae(p_e) with the
 * same Dirichlet treatment as the assembled operator. Needs the pattern, not the matrix. */
int ptb_set_operator_mode(ptb_ctx* ctx, int mode);
/* The `action` seam of cg.h:38-39 on its own: y = A p (halo update of p included); p_host has
 * (n_owned+n_ghost)*bs entries, y_host n_owned*bs. */
int ptb_apply_operator(ptb_ctx* ctx, const double* p_host, double* y_host);

/* ---- data in / out ------------------------------------------------------------------------ */
int ptb_set_rhs(ptb_ctx* ctx, const double* b_owned);
int ptb_set_initial_guess(ptb_ctx* ctx, const double* x_local); /* NULL resets to zero */
int ptb_get_matrix_values(ptb_ctx* ctx, double* vals);          /* [nnz*bs*bs], CSR order */
int ptb_get_diagonal_inverse(ptb_ctx* ctx, double* dinv);       /* [n_owned*bs] */
int ptb_get_rhs(ptb_ctx* ctx, double* b_owned);                 /* [n_owned*bs] */
int ptb_get_solution(ptb_ctx* ctx, double* x_local);            /* [(n_owned+n_ghost)*bs] */
/* ||u||_2 over owned entries, summed over ranks (la::norm, main.cpp:229). */
int ptb_solution_norm(ptb_ctx* ctx, double* norm);

/* ---- integer side, host only (no GPU needed) ---------------------------------------------- */
/* The cell -> CSR slot map: slot[(c*nd+i)*nd+j] = CSR position of (dofmap[c][i], dofmap[c][j]),
 * -1 for rows >= n_owned. */
int ptb_build_cell_slot_map(int64_t n_cells, int nd, const int32_t* dofmap, int32_t n_owned,
                            const int64_t* rowptr, const int32_t* cols, int64_t* slot);
/* The compressed form the kernels use, copied back from the context: pair list (dof ->
 * cell*nd+local index, ascending) and in-row offsets [n_pairs*nd]. Pass NULL to query sizes. */
int ptb_get_slot_offsets(ptb_ctx* ctx, int64_t* n_pairs, int64_t* pair_ptr, uint32_t* pairs,
                         uint16_t* offsets);
/* ---- instrumentation -------------------------------------------------------------------- */
/* Device time (CUDA events on the launching stream) of the last call of a stage, in ms. */
double ptb_stage_ms(const ptb_ctx* ctx, int stage);
/* Average device time in ms (CUDA events on the launching stream, after one warm-up launch) of
 * `reps` back-to-back launches of one hot kernel on the resident data. Scratches the solver
 * state (the next ptb_cg_solve re-initialises it); used by bench.py for the roofline numbers. */
enum { PTB_KERNEL_SPMV = 0, PTB_KERNEL_CG_UPDATE = 1, PTB_KERNEL_CG_DIRECTION = 2,
       PTB_KERNEL_ASSEMBLE_MATRIX = 3, PTB_KERNEL_ASSEMBLE_VECTOR = 4 };
int ptb_time_kernel(ptb_ctx* ctx, int which, int reps, double* ms_avg);
/* Kernels launched by this context so far. */
int64_t ptb_launch_count(const ptb_ctx* ctx);
/* Block entries (SELL padding included) the operator kernels actually stream per application:
 * the full pattern, or the zero-compacted copy when PTB_SPMV_COMPACT=1 built one. */
int64_t ptb_spmv_stored_entries(const ptb_ctx* ctx);
/* Fraction of the SpMV's column indices stored explicitly (the rest are one warp-uniform
 * delta per 32 rows; scalar matrices only, 1.0 otherwise). */
double ptb_cols_explicit_fraction(const ptb_ctx* ctx);
/* Bytes of device memory held by the context. */
int64_t ptb_device_bytes(const ptb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PTB200_H */
