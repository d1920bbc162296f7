// Host-only inspection entry points of libptb200.so (include/ptb200_debug.h: test hooks, not part of
// the drop-in boundary; ptb_build_cell_slot_map of include/ptb200.h "integer side, host only"):
// the integer structures the kernels read, rebuilt from their inputs without a GPU, for the CPU tests.
#include "abi_util.h"
#include "layout.h"
#include <algorithm>

using namespace ptb;
using ptb::abi::guarded;
using ptb::abi::need;

extern "C" {

int ptb_build_cell_slot_map(int64_t n_cells, int nd, const int32_t* dofmap, int32_t n_owned,
                            const int64_t* rowptr, const int32_t* cols, int64_t* slot)
{
  return guarded(nullptr, [&] {
    need(dofmap && rowptr && cols && slot, "ptb_build_cell_slot_map: NULL argument");
    build_cell_slot_map(dofmap, n_cells, nd, n_owned, rowptr, cols, slot);
  });
}

int ptb_debug_layout_roundtrip(int32_t n_rows, int64_t n_cols, const int64_t* rowptr,
                               const int32_t* cols, int32_t* cols_out, double* explicit_fraction)
{
  return guarded(nullptr, [&] {
    need(rowptr && cols && cols_out, "ptb_debug_layout_roundtrip: NULL argument");
    RowAdjacency adj;
    adj.ptr.assign(static_cast<std::size_t>(n_rows) + 1, 0);
    std::vector<std::uint16_t> so;
    SellLayout L;
    build_sell_layout(n_rows, 4, rowptr, cols, adj, so, 0, L);
    compress_columns(n_rows, n_cols, rowptr, L);
    for (std::int32_t s = 0; s < L.n_slices; ++s)
    {
      const std::int64_t mo = L.mat_off[s], w = (L.mat_off[s + 1] - mo) / 32;
      for (int lane = 0; lane < 32; ++lane)
      {
        const std::int32_t r = 32 * s + lane;
        if (r >= n_rows)
          continue;
        std::int64_t j = 0;
        for (std::int64_t k = 0; k < w; ++k)
        {
          const std::int32_t d = L.cdelta[mo / 32 + k];
          const std::int32_t cidx = d != CDELTA_EXPLICIT ? r + d : L.colsx[L.xoff[s] + (j++) * 32 + lane];
          if (k < rowptr[r + 1] - rowptr[r])
            cols_out[rowptr[r] + k] = cidx;
          else
            need(cidx >= 0 && cidx < n_cols, "layout: padding column out of range");
        }
      }
    }
    if (explicit_fraction)
      *explicit_fraction = L.cols.empty() ? 0.0 : static_cast<double>(L.colsx.size()) / L.cols.size();
  });
}

int ptb_debug_star_walk(int64_t n_cells, const int32_t* dofmap, int32_t n_owned,
                        const int64_t* rowptr, const int32_t* cols, uint32_t* walk_out,
                        double* loads_per_step)
{
  return guarded(nullptr, [&] {
    need(dofmap && rowptr && cols && walk_out, "ptb_debug_star_walk: NULL argument");
    RowAdjacency adj;
    std::vector<std::uint16_t> so;
    build_row_adjacency(dofmap, n_cells, 4, n_owned, adj);
    const std::int64_t max_so = build_slot_offsets(dofmap, 4, n_owned, adj, rowptr, cols, so);
    need(max_so >= 0 && max_so < 255, "ptb_debug_star_walk: pattern does not cover the cells / row too long");
    SellLayout L;
    build_sell_layout(n_owned, 4, rowptr, cols, adj, so, max_so, L);
    const WalkStats st = build_walk(n_owned, adj, so, L);
    for (std::int32_t r = 0; r < n_owned; ++r)
      for (std::int64_t k = 0; k < adj.ptr[r + 1] - adj.ptr[r]; ++k)
        walk_out[adj.ptr[r] + k] = L.walk[L.adj_off[r >> 5] + k * 32 + (r & 31)];
    if (loads_per_step)
      *loads_per_step = st.steps ? static_cast<double>(st.loads) / st.steps : 0.0;
  });
}

int ptb_debug_star_walk_single(int64_t n_cells, const int32_t* dofmap, int32_t n_owned,
                               const int64_t* rowptr, const int32_t* cols, int64_t* step_ptr,
                               uint32_t* words)
{
  return guarded(nullptr, [&] {
    need(dofmap && rowptr && cols && step_ptr, "ptb_debug_star_walk_single: NULL argument");
    RowAdjacency adj;
    std::vector<std::uint16_t> so;
    build_row_adjacency(dofmap, n_cells, 4, n_owned, adj);
    const std::int64_t max_so = build_slot_offsets(dofmap, 4, n_owned, adj, rowptr, cols, so);
    need(max_so >= 0 && max_so < 255, "ptb_debug_star_walk_single: pattern does not cover the cells / row too long");
    SellLayout L;
    build_sell_layout(n_owned, 4, rowptr, cols, adj, so, max_so, L);
    build_walk(n_owned, adj, so, L);
    build_walk_single(n_owned, adj, L);
    step_ptr[0] = 0;
    for (std::int32_t r = 0; r < n_owned; ++r)
    {
      const std::int64_t base = L.walk1_off[r >> 5] + (r & 31);
      const std::int64_t w1 = (L.walk1_off[(r >> 5) + 1] - L.walk1_off[r >> 5]) / 32;
      std::int64_t n = 0;
      while (n < w1 && L.walk1[base + n * 32] != ADJ_INVALID)
        ++n;
      if (words)
        for (std::int64_t k = 0; k < n; ++k)
          words[step_ptr[r] + k] = L.walk1[base + k * 32];
      step_ptr[r + 1] = step_ptr[r] + n;
    }
  });
}

int ptb_debug_p1_layout(int64_t n_cells, const int32_t* dofmap, int32_t n_owned,
                        const int64_t* rowptr, const int32_t* cols, int* max_w, int64_t* mat_off,
                        int64_t* adj_off, int64_t* walk1_off, int32_t* cols_sell, uint32_t* walk,
                        uint32_t* walk1, uint32_t* adjrot)
{
  return guarded(nullptr, [&] {
    need(dofmap && rowptr && cols && mat_off && adj_off && walk1_off, "ptb_debug_p1_layout: NULL argument");
    RowAdjacency adj;
    std::vector<std::uint16_t> so;
    build_row_adjacency(dofmap, n_cells, 4, n_owned, adj);
    const std::int64_t max_so = build_slot_offsets(dofmap, 4, n_owned, adj, rowptr, cols, so);
    need(max_so >= 0 && max_so < 255, "ptb_debug_p1_layout: pattern does not cover the cells / row too long");
    SellLayout L;
    build_sell_layout(n_owned, 4, rowptr, cols, adj, so, max_so, L);
    build_walk(n_owned, adj, so, L);
    build_walk_single(n_owned, adj, L);
    if (max_w)
      *max_w = L.max_w;
    std::copy(L.mat_off.begin(), L.mat_off.end(), mat_off);
    std::copy(L.adj_off.begin(), L.adj_off.end(), adj_off);
    std::copy(L.walk1_off.begin(), L.walk1_off.end(), walk1_off);
    if (cols_sell)
      std::copy(L.cols.begin(), L.cols.end(), cols_sell);
    if (walk)
      std::copy(L.walk.begin(), L.walk.end(), walk);
    if (walk1)
      std::copy(L.walk1.begin(), L.walk1.end(), walk1);
    if (adjrot)
      std::copy(L.adjrot.begin(), L.adjrot.end(), adjrot);
  });
}

int ptb_debug_p1_rings(int64_t n_cells, const int32_t* dofmap, int32_t n_owned, const int64_t* rowptr,
                       const int32_t* cols, int64_t* ring_off, uint8_t* ring_ns, uint32_t* ring)
{
  return guarded(nullptr, [&] {
    need(dofmap && rowptr && cols && ring_off && ring_ns, "ptb_debug_p1_rings: NULL argument");
    RowAdjacency adj;
    std::vector<std::uint16_t> so;
    build_row_adjacency(dofmap, n_cells, 4, n_owned, adj);
    const std::int64_t max_so = build_slot_offsets(dofmap, 4, n_owned, adj, rowptr, cols, so);
    need(max_so >= 0 && max_so < 127, "ptb_debug_p1_rings: pattern does not cover the cells / row too long");
    SellLayout L;
    build_sell_layout(n_owned, 4, rowptr, cols, adj, so, max_so, L);
    build_rings(n_owned, rowptr, adj, so, L);
    need(!L.ring_off.empty(), "ptb_debug_p1_rings: no rings for this pattern");
    std::copy(L.ring_off.begin(), L.ring_off.end(), ring_off);
    std::copy(L.ring_ns.begin(), L.ring_ns.end(), ring_ns);
    if (ring)
      std::copy(L.ring.begin(), L.ring.end(), ring);
  });
}

int ptb_debug_compressed_columns(int32_t n_rows, int64_t n_cols, const int64_t* rowptr,
                                 const int32_t* cols, int32_t* cdelta, int64_t* xoff,
                                 int32_t* colsx)
{
  return guarded(nullptr, [&] {
    need(rowptr && cols && cdelta && xoff, "ptb_debug_compressed_columns: NULL argument");
    RowAdjacency adj;
    adj.ptr.assign(static_cast<std::size_t>(n_rows) + 1, 0);
    std::vector<std::uint16_t> so;
    SellLayout L;
    build_sell_layout(n_rows, 4, rowptr, cols, adj, so, 0, L);
    compress_columns(n_rows, n_cols, rowptr, L);
    std::copy(L.cdelta.begin(), L.cdelta.end(), cdelta);
    std::copy(L.xoff.begin(), L.xoff.end(), xoff);
    if (colsx)
      std::copy(L.colsx.begin(), L.colsx.end(), colsx);
  });
}

int ptb_debug_balance_plan(int32_t n_slices, const int64_t* mat_off, const int32_t* order, int32_t n_interior,
                           int grid, int npull, int32_t* ounit, int32_t* begin, int32_t* n_begin)
{
  return guarded(nullptr, [&] {
    need(mat_off && order && ounit && begin && n_begin, "ptb_debug_balance_plan: NULL argument");
    std::vector<std::int32_t> ou, bg;
    const int longest = build_balance_plan(mat_off, order, n_slices, n_interior, grid, npull, ou, bg);
    std::copy(ou.begin(), ou.end(), ounit);
    std::copy(bg.begin(), bg.end(), begin);
    *n_begin = static_cast<std::int32_t>(bg.size());
    (void)longest;
  });
}

int ptb_debug_slice_order(int32_t n_rows, const int64_t* rowptr, const int32_t* cols,
                          int32_t* order, int32_t* n_interior)
{
  return guarded(nullptr, [&] {
    need(rowptr && cols && order && n_interior, "ptb_debug_slice_order: NULL argument");
    RowAdjacency adj;
    adj.ptr.assign(static_cast<std::size_t>(n_rows) + 1, 0);
    std::vector<std::uint16_t> so;
    SellLayout L;
    build_sell_layout(n_rows, 4, rowptr, cols, adj, so, 0, L);
    std::vector<std::int32_t> ord;
    build_slice_order(L, n_rows, 8, false, ord, *n_interior);
    std::copy(ord.begin(), ord.end(), order);
  });
}

int ptb_debug_pk_layout(int64_t n_cells, int nd, const int32_t* dofmap, int32_t n_owned,
                        const int64_t* rowptr, const int32_t* cols, int* info, int64_t* mat_off,
                        int64_t* adj_off, int32_t* bin_off, int* bin_w, int32_t* cols_sell,
                        uint32_t* adj, uint32_t* adjso, int32_t* bin_slices)
{
  return guarded(nullptr, [&] {
    need(dofmap && rowptr && cols && info && mat_off && adj_off && bin_off && bin_w,
         "ptb_debug_pk_layout: NULL argument");
    RowAdjacency a;
    std::vector<std::uint16_t> so;
    build_row_adjacency(dofmap, n_cells, nd, n_owned, a);
    const std::int64_t max_so = build_slot_offsets(dofmap, nd, n_owned, a, rowptr, cols, so);
    need(max_so >= 0, "ptb_debug_pk_layout: pattern does not cover the cells");
    SellLayout L;
    build_sell_layout(n_owned, nd, rowptr, cols, a, so, max_so, L);
    std::vector<std::int32_t> list, off;
    std::vector<int> width;
    build_width_bins(L, list, off, width);
    need(width.size() <= 16, "ptb_debug_pk_layout: more than 16 bins");
    info[0] = L.max_w, info[1] = L.so_bits, info[2] = L.so_words, info[3] = static_cast<int>(width.size());
    std::copy(L.mat_off.begin(), L.mat_off.end(), mat_off);
    std::copy(L.adj_off.begin(), L.adj_off.end(), adj_off);
    std::copy(off.begin(), off.end(), bin_off);
    std::copy(width.begin(), width.end(), bin_w);
    if (cols_sell)
      std::copy(L.cols.begin(), L.cols.end(), cols_sell);
    if (adj)
      std::copy(L.adj.begin(), L.adj.end(), adj);
    if (adjso)
      std::copy(L.adjso.begin(), L.adjso.end(), adjso);
    if (bin_slices)
      std::copy(list.begin(), list.end(), bin_slices);
  });
}

int ptb_debug_facet_rows(int64_t n_facets, const int32_t* cells, const int32_t* local_facets,
                         const int32_t* dofmap, int nd, int order, int32_t n_rows,
                         int32_t* n_frows, int32_t* n_ent, int32_t* row_ids, int32_t* row_ptr,
                         int32_t* ent)
{
  return guarded(nullptr, [&] {
    need(cells && local_facets && dofmap && n_frows && n_ent && row_ids && row_ptr && ent,
         "ptb_debug_facet_rows: NULL argument");
    std::vector<std::int32_t> ids, ptr, e;
    build_facet_rows(n_facets, cells, local_facets, dofmap, nd, order, n_rows, ids, ptr, e);
    need(e.size() <= static_cast<std::size_t>(20) * n_facets, "ptb_debug_facet_rows: capacity");
    *n_frows = static_cast<std::int32_t>(ids.size());
    *n_ent = static_cast<std::int32_t>(e.size() / 2);
    std::copy(ids.begin(), ids.end(), row_ids);
    std::copy(ptr.begin(), ptr.end(), row_ptr);
    std::copy(e.begin(), e.end(), ent);
  });
}

int ptb_debug_facet_rows_gathered(int64_t n_facets, const int32_t* cells, const int32_t* local_facets,
                                  const int32_t* gathered, int nd, int order, int32_t n_rows,
                                  int32_t* n_frows, int32_t* n_ent, int32_t* row_ids, int32_t* row_ptr,
                                  int32_t* ent)
{
  return guarded(nullptr, [&] {
    need(cells && local_facets && gathered && n_frows && n_ent && row_ids && row_ptr && ent,
         "ptb_debug_facet_rows_gathered: NULL argument");
    std::vector<std::int32_t> ids, ptr, e;
    build_facet_rows_gathered(n_facets, cells, local_facets, gathered, nd, order, n_rows, ids, ptr, e);
    need(e.size() <= static_cast<std::size_t>(20) * n_facets, "ptb_debug_facet_rows_gathered: capacity");
    *n_frows = static_cast<std::int32_t>(ids.size());
    *n_ent = static_cast<std::int32_t>(e.size() / 2);
    std::copy(ids.begin(), ids.end(), row_ids);
    std::copy(ptr.begin(), ptr.end(), row_ptr);
    std::copy(e.begin(), e.end(), ent);
  });
}

} // extern "C"
