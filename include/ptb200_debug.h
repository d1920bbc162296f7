/* ptb200_debug.h -- test hooks of libptb200.so. NOT part of the drop-in boundary (include/ptb200.h):
 * nothing here replaces a reference interface. These entry points expose the integer layouts the
 * kernels read (SELL-32 arrays, star walks, compressed columns, slice order, facet row lists) so
 * that tests/ can compare them bit for bit with the numpy restatements and run the kernel sources
 * on the host (tests/emu). A maintainer binding the library does not need this header.
 */
#ifndef PTB200_DEBUG_H
#define PTB200_DEBUG_H
#include "ptb200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* The P1 assembly maps as they sit on the device, downloaded (needs the GPU): adj_off
 * [n_slices + 1], the rotated slot words and the star-walk words [adj_off[n_slices]] (layout:
 * DESIGN.md section 3). Built on the host by default, by the setup kernels with PTB_GPU_SETUP=1
 * (*built_on_device reports which); tests compare them with ptb_debug_p1_layout bit for bit.
 * adjrot / walk may be NULL; *have_walk = 0 when the context holds no walk. */
int ptb_get_p1_maps(ptb_ctx* ctx, int64_t* adj_off, uint32_t* adjrot, uint32_t* walk,
                    int* have_walk, int* built_on_device);

/* The edge rings of the column-major elasticity P1 kernel as they sit on the device (needs the GPU):
 * ring_off [n_slices + 1], ring_ns [mat_off[n_slices] / 32], ring [ring_off[n_slices]] (layout.h
 * SellLayout::ring); any of the three may be NULL; *have_rings = 0 when the context holds none
 * (scalar problem, PTB_ASM_RING=0, rows longer than 127 columns). Tests compare them with
 * ptb_debug_p1_rings bit for bit for host-built and device-built (PTB_GPU_SETUP=1) maps. */
int ptb_get_p1_rings(ptb_ctx* ctx, int64_t* ring_off, uint8_t* ring_ns, uint32_t* ring, int* have_rings);

/* Round trip of the device matrix layout on the host (no GPU): builds the SELL-32 layout and the
 * compressed column indices from a CSR pattern, decodes them again into cols_out (CSR order) and
 * reports the fraction of indices that stayed explicit. Used by the CPU tests. */
int ptb_debug_layout_roundtrip(int32_t n_rows, int64_t n_cols, const int64_t* rowptr,
                               const int32_t* cols, int32_t* cols_out, double* explicit_fraction);

/* The P1 star walk the assembly kernels follow (host only, no GPU): for every owned row the
 * step words in walk order, walk_out[pair_ptr[r] + k] for step k of row r (pair_ptr = dof -> cell
 * adjacency offsets, as ptb_get_slot_offsets returns them). Byte p < 3 of a word = in-row offset
 * of the vertex held in register position p after the step, byte 3 = mask of positions loaded in
 * the step. loads_per_step (optional) receives the average number of vertices loaded per step. */
int ptb_debug_star_walk(int64_t n_cells, const int32_t* dofmap, int32_t n_owned,
                        const int64_t* rowptr, const int32_t* cols, uint32_t* walk_out,
                        double* loads_per_step);

/* The single-reload form of the walk (one new vertex per step; SellLayout::walk1): step_ptr
 * [n_owned + 1] receives the per-row step offsets, words (may be NULL to query sizes) the step
 * words row by row. Step 0 of a row has the walk format; later steps: byte 0 = offset of the new
 * vertex, byte 1 = offset of the evicted vertex, bits 16-17 = register position (3 = none),
 * bit 18 = a cell is complete after the step. Host only. */
int ptb_debug_star_walk_single(int64_t n_cells, const int32_t* dofmap, int32_t n_owned,
                               const int64_t* rowptr, const int32_t* cols, int64_t* step_ptr,
                               uint32_t* words);

/* The SELL-32 arrays the P1 walk kernels read (host only): offsets [ceil(n_owned/32) + 1] first
 * (pass NULL for the data arrays), then the data: padded columns, walk words, single-reload walk
 * words, rotated cell words (adjrot, indexed like walk), each in device order (offset[s] + k*32 + lane). Used by tests/emu, which runs the
 * kernel sources on the host. */
int ptb_debug_p1_layout(int64_t n_cells, const int32_t* dofmap, int32_t n_owned,
                        const int64_t* rowptr, const int32_t* cols, int* max_w, int64_t* mat_off,
                        int64_t* adj_off, int64_t* walk1_off, int32_t* cols_sell, uint32_t* walk,
                        uint32_t* walk1, uint32_t* adjrot);

/* The edge rings of the column-major elasticity P1 kernel (layout.h build_rings, csrc/assemble_ring.cu),
 * host only: ring_off [ceil(n_owned/32) + 1] (in words), ring_ns [mat_off[S]/32] (chain bytes per
 * lane of every (slice, column)); ring [ring_off[S]] may be NULL on the first call. */
int ptb_debug_p1_rings(int64_t n_cells, const int32_t* dofmap, int32_t n_owned, const int64_t* rowptr,
                       const int32_t* cols, int64_t* ring_off, uint8_t* ring_ns, uint32_t* ring);

/* The compressed column indices of the scalar SpMV (layout.h: cdelta / xoff / colsx), host only:
 * cdelta [mat_off[S]/32], xoff [S + 1]; colsx [xoff[S]] may be NULL on the first call. */
int ptb_debug_compressed_columns(int32_t n_rows, int64_t n_cols, const int64_t* rowptr,
                                 const int32_t* cols, int32_t* cdelta, int64_t* xoff,
                                 int32_t* colsx);

/* The balanced split of the operator kernels for small problems (layout.h build_balance_plan), host
 * only: ounit [n_slices + 1]; begin holds at most grid + 2 entries, *n_begin receives the count. */
int ptb_debug_balance_plan(int32_t n_slices, const int64_t* mat_off, const int32_t* order, int32_t n_interior,
                           int grid, int npull, int32_t* ounit, int32_t* begin, int32_t* n_begin);

/* The slice visiting order of the operator kernels (layout.h build_slice_order; slices without
 * ghost columns first), host only: order [ceil(n_rows/32)], *n_interior = number of leading
 * slices that read no ghost column. */
int ptb_debug_slice_order(int32_t n_rows, const int64_t* rowptr, const int32_t* cols,
                          int32_t* order, int32_t* n_interior);

/* The SELL-32 arrays of the P2/P3 assembly kernels and the row-length bins of the binned matrix
 * kernel (host only, used by tests/emu). info = {max_w, so_bits, so_words, n_bins}; offsets
 * [ceil(n_owned/32) + 1]; bin_off [17], bin_w [16]; data arrays may be NULL on the first call:
 * cols_sell [mat_off[S]], adj [adj_off[S]], adjso [adj_off[S] * so_words], bin_slices [S]. */
int ptb_debug_pk_layout(int64_t n_cells, int nd, const int32_t* dofmap, int32_t n_owned,
                        const int64_t* rowptr, const int32_t* cols, int* info, int64_t* mat_off,
                        int64_t* adj_off, int32_t* bin_off, int* bin_w, int32_t* cols_sell,
                        uint32_t* adj, uint32_t* adjso, int32_t* bin_slices);

/* The boundary-facet gather lists of assemble_vector (layout.h build_facet_rows), host only.
 * Capacities: row_ids [n_rows], row_ptr [n_rows + 1], ent [2 * 10 * n_facets]; *n_frows and
 * *n_ent receive the used lengths (ent holds *n_ent (cell, local_facet*nd + li) pairs). */
int ptb_debug_facet_rows(int64_t n_facets, const int32_t* cells, const int32_t* local_facets,
                         const int32_t* dofmap, int nd, int order, int32_t n_rows,
                         int32_t* n_frows, int32_t* n_ent, int32_t* row_ids, int32_t* row_ptr,
                         int32_t* ent);
/* The same lists from the dofmap rows of the facets' cells only, gathered [n_facets * nd] with
 * gathered[k*nd + j] = dofmap[cells[k]*nd + j] (the route ptb_set_exterior_facets takes when the
 * dofmap was generated on the device). */
int ptb_debug_facet_rows_gathered(int64_t n_facets, const int32_t* cells, const int32_t* local_facets,
                         const int32_t* gathered, int nd, int order, int32_t n_rows,
                         int32_t* n_frows, int32_t* n_ent, int32_t* row_ids, int32_t* row_ptr,
                         int32_t* ent);

#ifdef __cplusplus
}
#endif
#endif /* PTB200_DEBUG_H */
