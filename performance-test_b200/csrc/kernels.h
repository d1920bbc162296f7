// Kernel argument blocks and launchers shared by the .cu files.
#pragma once
#include <vector>
// PTB_HOST_EMU exists for tests/emu only (the g++ harness that executes kernel *sources* on the
// host to check their indexing). It must never reach a device build: there is no CPU path in the
// product, and a library built with it would be one.
#if defined(PTB_HOST_EMU) && defined(__CUDACC__)
#error "PTB_HOST_EMU is a test-harness switch (tests/emu); it is not valid in an nvcc build"
#endif
#include "ctx.h"
#include "peer.h"

namespace ptb
{

#define ADJ_INVALID_DEV 0xFFFFFFFFu

struct MatrixArgs
{
  std::int32_t n_rows, n_slices;
  int so_bits, so_words;
  const double* xyz;
  const std::int32_t* x_dofmap;
  const std::int32_t* dofmap;
  const std::uint8_t* bc;
  const std::int64_t* rowptr;
  const std::int64_t* mat_off;
  const std::int64_t* adj_off;
  const std::int32_t* cols;
  const std::uint32_t* adj;
  const std::uint32_t* adjso;
  const std::uint32_t* adjrot;
  const double* xdof;
  int max_w;
  double* vals;
  double* dinv;
  // P2/P3: geometry factors per cell (assemble_pk.cu cell_geometry_pk), 8 doubles per cell:
  // G00 G01 G02 G11 G12 G22 |det| 0; nullptr = the kernels compute them per (row, cell) pair
  const double* cell_g = nullptr;
};

struct VectorArgs
{
  std::int32_t n_rows, n_slices;
  const double* xyz;
  const std::int32_t* x_dofmap;
  const std::int32_t* dofmap;
  const std::uint8_t* bc;
  const std::int64_t* adj_off;
  const std::uint32_t* adj;
  const std::uint32_t* adjrot;
  const double* xdof;
  const std::int64_t* mat_off;
  const std::int32_t* cols;
  int max_w;
  const double* f;
  double* b;
};

struct FacetArgs
{
  std::int32_t n_frows;
  const double* xyz;
  const std::int32_t* x_dofmap;
  const std::int32_t* dofmap;
  const std::uint8_t* bc;
  const std::int32_t* frow_ids;
  const std::int32_t* frow_ptr;
  const std::int32_t* fent;
  const double* g;
  double* b;
};

/// SELL-32 matrix view for the SpMV.
struct SpmvArgs
{
  std::int32_t n_rows, n_slices;
  const std::int64_t* mat_off;
  const std::int32_t* cols;
  const double* vals;
  const std::int32_t* cdelta; // scalar matrices: compressed columns (nullptr: use cols)
  const std::int32_t* colsx;
  const std::int64_t* xoff;
  // L2 plan (sync_ops.cuh l2_policy): l2_mode 0 = no hints; 1 = slices stored below pin_entries
  // (mat_off units) are loaded evict_last, the rest evict_first
  int l2_mode;
  std::int64_t pin_entries;
  // Balanced work split of small problems (cg.cu spmv_cta_balanced), nullptr = one slice per warp
  // step: ounit [n_slices + 1] = stored entries per row (k-steps) before position i of the slice
  // order; bal_begin = the CTAs' runs of positions (fused halo: npull + 1 entries for the pullers'
  // ghost-reading slices, then the workers' entries; otherwise gridDim.x + 1 entries)
  const std::int32_t* ounit;
  const std::int32_t* bal_begin;
};

/// The operator view the SpMV kernels read: the compacted copy when one is current (compact.cu).
inline SpmvArgs spmv_args(const ptb_ctx* c)
{
  if (c->have_compact)
    return SpmvArgs{c->n_owned, c->n_slices, c->mat_off_z.p, c->cols.p, c->vals_z.p,
                    c->cdelta_z.p, c->colsx_z.p, c->xoff_z.p, c->l2_mode, c->l2_pin_entries,
                    nullptr, nullptr};
  return SpmvArgs{c->n_owned, c->n_slices, c->mat_off.p, c->cols.p, c->vals.p,
                  c->cdelta.p, c->colsx.p, c->xoff.p, c->l2_mode, c->l2_pin_entries,
                  nullptr, nullptr};
}
/// Build the zero-column-compacted copy of the assembled scalar operator (no-op for bs = 3).
void compact_operator(ptb_ctx* c);

/// Device-side construction of the P1 assembly maps (setup.cu, opt-in PTB_GPU_SETUP=1): fills
/// c->adj_off, c->adjrot and, if want_walk, c->walk from the uploaded dofmap, rowptr, mat_off and
/// padded columns. Returns false when the pattern cannot be expressed (the caller builds on the host).
/// want_rings: also c->ring, c->ring_off, c->ring_ns (edge rings of assemble_ring.cu; rows of at most 127 columns).
bool gpu_setup_p1(ptb_ctx* c, bool want_walk, bool want_rings, int* max_wa);
/// The same for P2/P3: c->adj_off, c->adj, c->adjso with 8-bit offsets (rows of at most 256 columns).
bool gpu_setup_pk(ptb_ctx* c, int* max_wa);
/// The sparsity pattern of the owned rows built on the device from the uploaded dofmap (setup.cu) and
/// downloaded; false when a row is too long for the device build.
/// rp / cl keep the device copies (CSR).
bool gpu_build_pattern(ptb_ctx* c, std::vector<std::int64_t>& rowptr, std::vector<std::int32_t>& cols,
                       DevBuf<std::int64_t>& rp, DevBuf<std::int32_t>& cl);
/// Column side of ptb_set_pattern (SELL-32 offsets and padded columns, column compression, slice
/// order) from a CSR pattern on the device; takes rp over as c->rowptr. h_mat_off = host copy of mat_off.
void gpu_setup_columns(ptb_ctx* c, DevBuf<std::int64_t>& rp, const DevBuf<std::int32_t>& cl,
                       std::vector<std::int64_t>& h_mat_off);

/// The local z-slab of the unit-cube Kuhn mesh and its Lagrange dofmap (order 1..3) generated on the
/// device (box.cu).
void gpu_create_box(ptb_ctx* c, int order, std::int64_t nx, std::int64_t ny, std::int64_t nz, int rank, int nranks);
/// out[k*nd + j] = dofmap[cells[k]*nd + j] for the n listed cells (box.cu).
void launch_gather_dofmap_rows(ptb_ctx* c, std::int64_t n, const std::int32_t* cells, std::int32_t* out);
/// Device-side problem data (problem_data.cu): Dirichlet markers from the reference's facet predicate
/// + facet closure into c->bc; the source terms at the dof coordinates X (stride 3 or 4) into c->f, c->g.
void launch_locate_bc(ptb_ctx* c);
void launch_interpolate_source(ptb_ctx* c, const double* X, int stride);

void launch_assemble_matrix(ptb_ctx* c, const MatrixArgs& A);
void launch_assemble_vector(ptb_ctx* c, const VectorArgs& A, const FacetArgs& F);
/// Star-walk variant (assemble_walk.cu); returns false when it does not apply (no walk uploaded,
/// not scalar P1, rows longer than 32 columns) and the caller falls through to the default kernel.
bool launch_assemble_matrix_walk(ptb_ctx* c, const MatrixArgs& A);
/// Column-major elasticity P1 kernel along the edge rings (assemble_ring.cu); false when the rings are
/// not on the device (PTB_ASM_RING=0, device-built maps, rows longer than 127 columns).
bool launch_assemble_matrix_ring(ptb_ctx* c, const MatrixArgs& A);
/// Direct-gather walk kernel of the P1 cell vector (assemble_gwalk.cu); same convention: false
/// when the single-reload walk is not on the device (PTB_VEC_GWALK=0, device-built maps).
bool launch_assemble_vector_gwalk(ptb_ctx* c, const VectorArgs& A);
void launch_assemble_matrix_pk(ptb_ctx* c, const MatrixArgs& A);
void launch_assemble_vector_pk(ptb_ctx* c, const VectorArgs& A, const FacetArgs& F);
/// Matrix-free y = A p (Poisson P1); py_out (device, optional) receives the local p.y.
void launch_action_matrix_free(ptb_ctx* c, const VectorArgs& A, const double* p, double* y,
                               double* py_out);
void launch_action_matrix_free_pk(ptb_ctx* c, const VectorArgs& A, const double* p, double* y,
                                  double* py_out);
void launch_reduce_partials(ptb_ctx* c, std::int64_t n, const double* partials, double* out);
void launch_sell_to_csr(ptb_ctx* c, double* out);
/// xdof[d] = xyz[dof_vertex[d]] for vertex dofs.
void launch_gather_xdof(ptb_ctx* c);
/// xyz[v][0..2] = xyz3[v][0..2] (pad the caller's 3-column geometry to 32-byte points).
void launch_pad_xyz(ptb_ctx* c);

// cg.cu
int cg_grid(const ptb_ctx* c);
/// y = A p on owned rows; if st != nullptr also st->py = p.y (local sum) and honours st->conv.
void launch_spmv(ptb_ctx* c, const double* p, double* y, CgState* st, unsigned int epoch = 0,
                 bool fused_halo = false);
/// r = b - y; p = dinv*r (owned); st: rnorm0 = rnorm = rr, rz_old = rz (local sums into rr, rz).
void launch_cg_init(ptb_ctx* c, const double* dinv, CgState* st, unsigned int epoch);
void launch_cg_finish_init(ptb_ctx* c, CgState* st, double rtol, unsigned int epoch);
/// x += alpha p; r -= alpha y; local sums r.r, r.z into cur->rr, cur->rz.
void launch_cg_update(ptb_ctx* c, const double* dinv, CgState* cur, unsigned int epoch_in,
                      unsigned int epoch_out);
/// beta, convergence test, bookkeeping into nxt; p = beta p + dinv r unless converged.
void launch_cg_direction(ptb_ctx* c, const double* dinv, const CgState* cur, CgState* nxt,
                         unsigned int epoch);
/// Iterations it0+1 .. it0+n_it of the CG loop in one cooperative kernel (cg.cu cg_loop); the
/// state records must have been initialised (launch_cg_finish_init). Returns false when the path
/// does not apply (the caller then queues the three kernels per iteration).
bool launch_cg_loop(ptb_ctx* c, const double* dinv, int it0, int n_it, unsigned int ebase,
                    unsigned int lbase, bool fused_halo);
/// Reduction epochs of the peer-memory all-reduce (never 0; see peer.cuh).
inline unsigned int next_red_epoch(ptb_ctx* c)
{
  if (++c->peer.red_epoch == 0)
    ++c->peer.red_epoch;
  return c->peer.red_epoch;
}
void launch_fill(ptb_ctx* c, double* v, std::int64_t n, double value);
void launch_pack(ptb_ctx* c, const double* v, const std::int32_t* idx, std::int64_t n, int bs,
                 double* out);
void launch_unpack(ptb_ctx* c, const double* in, const std::int32_t* idx, std::int64_t n, int bs,
                   double* v);
void launch_sqnorm(ptb_ctx* c, const double* v, std::int64_t n, double* out_dev);
/// Peer mode: publish st->py (computed outside spmv_sell) to the window all-reduce; else no-op.
void launch_publish_py(ptb_ctx* c, CgState* st, unsigned int epoch);

// peer.cu
PeerView peer_view(const ptb_ctx* c);
PeerHalo peer_halo(const ptb_ctx* c);

} // namespace ptb
