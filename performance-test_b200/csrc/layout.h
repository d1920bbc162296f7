// Host-side construction of the SELL-32 device layouts from the CSR pattern and the compressed
// cell -> slot map (see ctx.h for the layout).
#pragma once
#include "../common/intmaps.h"
#include <climits>
#include <cstdint>
#include <vector>

namespace ptb
{

constexpr std::uint32_t ADJ_INVALID = 0xFFFFFFFFu;

struct SellLayout
{
  std::int32_t n_slices = 0;
  int max_w = 0, max_wa = 0;
  int so_bits = 8, so_words = 1; // slot-offset packing
  std::vector<std::int64_t> mat_off, adj_off; // [n_slices + 1], in entries
  std::vector<std::int32_t> cols;             // padded with the row's first column (0 past n_rows)
  std::vector<std::uint32_t> adj, adjso;
  // P1 only (nd == 4, offsets < 255): the four in-row offsets of a pair rotated so that the
  // owner's own column comes first: byte t = offset of local vertex (li + t) & 3.
  std::vector<std::uint32_t> adjrot;
  // P1 only, same indexing as adjrot: the row's cells re-ordered into a *walk* over its vertex
  // star (build_walk). Word k of a row describes step k: bytes 0..2 = in-row offsets of the
  // three non-owner vertices held in register positions 0..2 after the step, byte 3 = mask of
  // the positions that were (re)loaded in this step (7 on the first cell). ADJ_INVALID = padding.
  std::vector<std::uint32_t> walk;
  // The same walk with exactly one vertex (re)loaded per step (build_walk_single), for kernels that
  // gather the new vertex straight from global memory a few steps ahead. Own SELL offsets walk1_off
  // (a cell that brings two or three new vertices takes two or three steps). Step 0 of a row keeps
  // the walk format (three offsets, mask 7). Later steps: byte 0 = in-row offset of the new vertex,
  // byte 1 = offset of the vertex it evicts (the accumulator to flush), bits 16-17 = register
  // position (3 = none), bit 18 = a cell is complete after this step. ADJ_INVALID = padding.
  std::vector<std::uint32_t> walk1;
  std::vector<std::int64_t> walk1_off; // [n_slices + 1], in entries
  int max_w1 = 0;
  // Column-index compression for the SpMV (scalar matrices): cdelta[mat_off[s]/32 + k] = d when
  // every row r of slice s has col_k = r + d (translation-invariant stencil), else CDELTA_EXPLICIT
  // and the 32 indices are stored in colsx at xoff[s] + j*32 + lane (j-th explicit k of the slice).
  std::vector<std::int32_t> cdelta, colsx;
  std::vector<std::int64_t> xoff; // [n_slices + 1]
  // P1 only (row length <= 127): *edge rings* (build_rings). For row i and stored column k (a
  // neighbour j != i) the cells that hold both i and j form a ring (interior edge) or fans
  // (boundary edge) around the edge (i, j); they are stored as a chain of vertices
  // v_0, v_1, ..., cell t = (i, j, v_{t-1}, v_t), one byte per vertex: bits 0-6 = in-row offset of
  // the vertex, bit 7 = restart (the vertex opens a chain: no cell is formed with its predecessor;
  // always set on byte 0). Padding = 0x80. ring_ns[mat_off[s]/32 + k] = bytes per lane of column k in
  // slice s (longest chain of the 32 rows; 0 where no row has a ring, e.g. the diagonal); the bytes
  // are packed four to a word, word q of column k of lane l at
  // ring[ring_off[s] + (sum_{k' < k} ceil(ring_ns[k'] / 4) + q) * 32 + l].
  std::vector<std::uint32_t> ring;
  std::vector<std::int64_t> ring_off; // [n_slices + 1], in words
  std::vector<std::uint8_t> ring_ns;  // [mat_off[n_slices] / 32]
};
constexpr std::int32_t CDELTA_EXPLICIT = INT32_MIN;

/// Build cdelta / colsx / xoff from the padded SELL columns. Padding entries (value 0) may take
/// any valid column, so they never break uniformity unless r + d leaves [0, n_cols).
void compress_columns(std::int32_t n_rows, std::int64_t n_cols, const std::int64_t* rowptr,
                      SellLayout& L);

void build_sell_layout(std::int32_t n_rows, int nd, const std::int64_t* rowptr,
                       const std::int32_t* cols, const RowAdjacency& adj,
                       const std::vector<std::uint16_t>& so, std::int64_t max_so,
                       SellLayout& L);

/// Star walk of every P1 row (requires L.adjrot, i.e. nd == 4 and offsets < 255). The cells of a
/// row are visited so that consecutive cells share as many vertices as possible (greedy: most
/// shared vertices with the current cell, ties to the earlier cell of the ascending list; the walk
/// starts at the row's first cell). Vertices that stay keep their register position; new vertices
/// take the freed positions in ascending order, in the cell's own rotation order (li+1, li+2, li+3),
/// so the walk depends on the mesh topology and the cell order only, never on dof labels.
/// On the Kuhn box every interior star (24 cells) is walked with one new vertex per step.
struct WalkStats
{
  std::int64_t steps = 0, loads = 0; // loads = vertices (re)loaded over all steps
};
WalkStats build_walk(std::int32_t n_rows, const RowAdjacency& adj,
                     const std::vector<std::uint16_t>& so, SellLayout& L);

/// Edge rings of every P1 row (see SellLayout::ring) for the column-major elasticity kernel
/// (assemble_ring.cu). The chains depend on the mesh topology and the ascending cell order only:
/// a chain starts at the lowest unvisited cell that has a vertex no other unvisited cell of the ring
/// shares (the end of a fan; the end vertex comes first), else at the lowest unvisited cell,
/// walking towards its lower face neighbour; it continues through the lowest unvisited cell that
/// holds the current vertex. Leaves L.ring empty when a row is longer than 127 columns.
/// Returns the number of chain bytes over all rows (cells + chains).
std::int64_t build_rings(std::int32_t n_rows, const std::int64_t* rowptr, const RowAdjacency& adj,
                         const std::vector<std::uint16_t>& so, SellLayout& L);

/// Derive walk1 / walk1_off from L.walk (see SellLayout::walk1).
void build_walk_single(std::int32_t n_rows, const RowAdjacency& adj, SellLayout& L);

/// Slices grouped by row-length class (SELL width <= 32, 64, 96, 128, 192, 256, ... max_w): list =
/// slice indices bin after bin (ascending inside a bin), off = [n_bins + 1] offsets into list,
/// width = accumulator width of each bin (the class bound, capped by max_w).
void build_width_bins(const SellLayout& L, std::vector<std::int32_t>& list,
                      std::vector<std::int32_t>& off, std::vector<int>& width);

/// Visiting order of the slices for the operator kernels. Slices without ghost columns come first
/// (n_interior of them), so the fused halo pull overlaps with them. Inside each class the order is
/// built from groups of `group` slices that reference each other's rows (breadth-first over the
/// slice graph from the lowest unvisited slice): the warps of one CTA work on one group at a time,
/// so the entries of p they gather overlap and are served by L1 instead of L2.
void build_slice_order(const SellLayout& L, std::int32_t n_rows, int group, bool cluster,
                       std::vector<std::int32_t>& order, std::int32_t& n_interior);

/// Boundary-facet gather lists: for every owned row touched by an exterior facet, the entries
/// (facet k, local dof li) with dofmap[cell_k][li] == row and li on the facet, ascending in k.
/// ent holds two ints per entry: cell, local_facet*nd + li.
void build_facet_rows(std::int64_t n_facets, const std::int32_t* cells,
                      const std::int32_t* local_facets, const std::int32_t* dofmap, int nd,
                      int order, std::int32_t n_rows, std::vector<std::int32_t>& row_ids,
                      std::vector<std::int32_t>& row_ptr, std::vector<std::int32_t>& ent);

/// The same lists from the dofmap rows of the facets' cells only (gathered[k*nd + j] =
/// dofmap[cells[k]*nd + j]): what a context whose dofmap lives on the device downloads instead of
/// the whole dofmap.
void build_facet_rows_gathered(std::int64_t n_facets, const std::int32_t* cells,
                               const std::int32_t* local_facets, const std::int32_t* gathered, int nd,
                               int order, std::int32_t n_rows, std::vector<std::int32_t>& row_ids,
                               std::vector<std::int32_t>& row_ptr, std::vector<std::int32_t>& ent);

/// Balanced split of the operator kernels for small problems (cg.cu spmv_cta_balanced): ounit
/// [S + 1] = k-steps (stored entries per row) before position i of the slice order; begin = the CTAs'
/// runs of positions, equal in k-steps up to one slice: with puller roles (npull >= 0) npull + 1
/// entries over the ghost-reading positions [n_interior, S) followed by grid - npull + 1 entries over
/// [0, n_interior); without roles grid + 1 entries over [0, S). Returns the longest run in slices.
int build_balance_plan(const std::int64_t* mat_off, const std::int32_t* order, std::int32_t n_slices,
                       std::int32_t n_interior, int grid, int npull, std::vector<std::int32_t>& ounit,
                       std::vector<std::int32_t>& begin);

} // namespace ptb
