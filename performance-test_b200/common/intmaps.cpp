#include "intmaps.h"
#include <algorithm>
#include <omp.h>

namespace ptb
{

void build_row_adjacency(const std::int32_t* dofmap, std::int64_t ncells, int nd,
                         std::int32_t n_rows, RowAdjacency& adj)
{
  adj.ptr.assign(static_cast<std::size_t>(n_rows) + 1, 0);
  std::int64_t* ptr = adj.ptr.data();
  const std::int64_t n = ncells * nd;

  // Count (row r is stored at ptr[r + 1]).
#pragma omp parallel for schedule(static)
  for (std::int64_t k = 0; k < n; ++k)
  {
    const std::int32_t d = dofmap[k];
    if (d < n_rows)
    {
#pragma omp atomic
      ptr[d + 1] += 1;
    }
  }
  for (std::int32_t r = 0; r < n_rows; ++r)
    ptr[r + 1] += ptr[r];

  adj.pairs.resize(static_cast<std::size_t>(ptr[n_rows]));
  std::uint32_t* pairs = adj.pairs.data();

  // Fill in arbitrary order, then sort each (short) row list: the result does not depend on the
  // thread count.
  std::vector<std::int64_t> cursor(ptr, ptr + n_rows);
  std::int64_t* cur = cursor.data();
#pragma omp parallel for schedule(static)
  for (std::int64_t k = 0; k < n; ++k)
  {
    const std::int32_t d = dofmap[k];
    if (d < n_rows)
    {
      std::int64_t pos;
#pragma omp atomic capture
      pos = cur[d]++;
      pairs[pos] = static_cast<std::uint32_t>(k);
    }
  }
#pragma omp parallel for schedule(dynamic, 4096)
  for (std::int32_t r = 0; r < n_rows; ++r)
    std::sort(pairs + ptr[r], pairs + ptr[r + 1]);
}

void build_pattern(const std::int32_t* dofmap, int nd, std::int32_t n_rows,
                   const RowAdjacency& adj, std::vector<std::int64_t>& rowptr,
                   std::vector<std::int32_t>& cols)
{
  rowptr.assign(static_cast<std::size_t>(n_rows) + 1, 0);
  const std::int64_t* aptr = adj.ptr.data();
  const std::uint32_t* pairs = adj.pairs.data();

  // Pass 1: row lengths. Pass 2: fill. The per-row work is a sort + unique of <= ncell_r * nd
  // candidates.
  for (int pass = 0; pass < 2; ++pass)
  {
    if (pass == 1)
    {
      for (std::int32_t r = 0; r < n_rows; ++r)
        rowptr[r + 1] += rowptr[r];
      cols.resize(static_cast<std::size_t>(rowptr[n_rows]));
    }
#pragma omp parallel
    {
      std::vector<std::int32_t> tmp;
#pragma omp for schedule(dynamic, 2048)
      for (std::int32_t r = 0; r < n_rows; ++r)
      {
        tmp.clear();
        for (std::int64_t k = aptr[r]; k < aptr[r + 1]; ++k)
        {
          const std::int64_t c = pairs[k] / static_cast<std::uint32_t>(nd);
          const std::int32_t* cd = dofmap + c * nd;
          tmp.insert(tmp.end(), cd, cd + nd);
        }
        std::sort(tmp.begin(), tmp.end());
        const auto last = std::unique(tmp.begin(), tmp.end());
        if (pass == 0)
          rowptr[r + 1] = last - tmp.begin();
        else
          std::copy(tmp.begin(), last, cols.begin() + rowptr[r]);
      }
    }
  }
}

std::int64_t build_slot_offsets(const std::int32_t* dofmap, int nd, std::int32_t n_rows,
                                const RowAdjacency& adj, const std::int64_t* rowptr,
                                const std::int32_t* cols, std::vector<std::uint16_t>& off)
{
  const std::int64_t* aptr = adj.ptr.data();
  const std::uint32_t* pairs = adj.pairs.data();
  off.resize(adj.pairs.size() * static_cast<std::size_t>(nd));
  std::int64_t maxoff = 0;
  bool bad = false;
#pragma omp parallel for schedule(dynamic, 2048) reduction(max : maxoff) reduction(|| : bad)
  for (std::int32_t r = 0; r < n_rows; ++r)
  {
    const std::int32_t* rc = cols + rowptr[r];
    const std::int64_t len = rowptr[r + 1] - rowptr[r];
    for (std::int64_t k = aptr[r]; k < aptr[r + 1]; ++k)
    {
      const std::int64_t c = pairs[k] / static_cast<std::uint32_t>(nd);
      const std::int32_t* cd = dofmap + c * nd;
      for (int j = 0; j < nd; ++j)
      {
        const std::int32_t* it = std::lower_bound(rc, rc + len, cd[j]);
        const std::int64_t o = it - rc;
        if (o >= len || *it != cd[j] || o > 65535)
          bad = true;
        off[k * nd + j] = static_cast<std::uint16_t>(o);
        maxoff = std::max(maxoff, o);
      }
    }
  }
  return bad ? -1 : maxoff;
}

void build_cell_slot_map(const std::int32_t* dofmap, std::int64_t ncells, int nd,
                         std::int32_t n_rows, const std::int64_t* rowptr,
                         const std::int32_t* cols, std::int64_t* slot)
{
#pragma omp parallel for schedule(static)
  for (std::int64_t c = 0; c < ncells; ++c)
  {
    const std::int32_t* cd = dofmap + c * nd;
    for (int i = 0; i < nd; ++i)
    {
      std::int64_t* s = slot + (c * nd + i) * nd;
      const std::int32_t r = cd[i];
      if (r >= n_rows)
      {
        for (int j = 0; j < nd; ++j)
          s[j] = -1;
        continue;
      }
      const std::int32_t* rc = cols + rowptr[r];
      const std::int64_t len = rowptr[r + 1] - rowptr[r];
      for (int j = 0; j < nd; ++j)
      {
        const std::int32_t* it = std::lower_bound(rc, rc + len, cd[j]);
        s[j] = (it < rc + len && *it == cd[j]) ? rowptr[r] + (it - rc) : -1;
      }
    }
  }
}

} // namespace ptb
