// Device-side construction of the P1 assembly maps (opt-in, PTB_GPU_SETUP=1) -- SURVEY 8f row 2:
// the cell -> CSR-slot map "built bit-exactly on the integer side", here on the device instead of
// the host (common/intmaps.cpp + layout.cpp build_sell_layout / build_walk). Everything is integer
// work; every array must equal the host build bit for bit (tests/emu runs these sources on the
// host against it).
//   count / scan / fill / sort   dof -> (cell, local index) pairs, ascending per row
//                                (build_row_adjacency: atomics place the pairs, the per-row sort
//                                makes the result independent of their order)
//   widths / scan                adj_off of the SELL-32 cell lists
//   rotated words                per (row, cell): binary search of the cell's four vertices in the
//                                row's column list, packed with the owner first (adjrot)
//   walk                         the greedy star walk of layout.cpp build_walk, one thread per row
//   sell cols / cdelta / colsx   the column side of ptb_set_pattern (layout.cpp build_sell_layout,
//   slice flags / order          compress_columns, build_slice_order without clustering): SELL-32
//                                offsets and padded columns, one delta per (slice, k) where the 32
//                                rows share col - row, explicit indices elsewhere, and the visiting
//                                order "slices without ghost columns first"
//   pattern count / scan / fill  the sparsity pattern itself (common/intmaps.cpp build_pattern; the
//                                reference builds it in fem::create_matrix, poisson_problem.cpp:122-123,
//                                inside the timed assembly stage): per owned row the ascending union of
//                                the dofs of its cells, any Lagrange order (ptb_build_pattern)
// With a caller-supplied pattern (ptb_set_pattern) the column side stays on the host: it only needs
// the CSR arrays, no adjacency. With ptb_build_pattern the pattern is already on the device and the
// column side is built there too (gpu_setup_columns), so no O(nnz) host loop is left for P1.
// Run on the B200 since round 2: every array bit-identical to the host build (GPU tests), set_problem
// 1.78 -> 0.74 s at 4 M DOFs (profiles/r02/setup_4M.json).
#include "kernels.h"
#include <climits>

namespace ptb
{
namespace
{

constexpr int SU_THREADS = 256;
constexpr int SU_MAX_CELLS = 64; // cells per row the walk kernel holds (Kuhn box: 24)

// pairs per row (row r is counted at cnt[r])
__global__ void setup_count(std::int64_t n_entries, const std::int32_t* __restrict__ dofmap,
                            std::int32_t n_rows, unsigned long long* __restrict__ cnt)
{
  const std::int64_t k = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (k >= n_entries)
    return;
  const std::int32_t d = dofmap[k];
  if (d < n_rows)
    atomicAdd(cnt + d, 1ull);
}

// Exclusive prefix sum of n values into out[0..n] (out[n] = total), one CTA (see compact.cu).
__global__ void __launch_bounds__(1024)
setup_scan(std::int64_t n, const unsigned long long* __restrict__ in, std::int64_t* __restrict__ out,
           std::int64_t scale)
{
  __shared__ std::int64_t part[1024];
  const std::int64_t chunk = (n + blockDim.x - 1) / blockDim.x;
  const std::int64_t lo = min(n, chunk * threadIdx.x), hi = min(n, lo + chunk);
  std::int64_t sum = 0;
  for (std::int64_t i = lo; i < hi; ++i)
    sum += static_cast<std::int64_t>(in[i]) * scale;
  part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    std::int64_t run = 0;
    for (unsigned t = 0; t < blockDim.x; ++t)
    {
      const std::int64_t v = part[t];
      part[t] = run;
      run += v;
    }
    out[n] = run;
  }
  __syncthreads();
  std::int64_t run = part[threadIdx.x];
  for (std::int64_t i = lo; i < hi; ++i)
  {
    out[i] = run;
    run += static_cast<std::int64_t>(in[i]) * scale;
  }
}

// Long inputs: the same prefix sum in three passes over tiles of SC_TILE elements -- tile sums, a
// scan of the tile sums (setup_scan above), and the tiles themselves: every thread owns SC_ITEMS
// consecutive elements (a warp reads 2 KB contiguously), the thread sums are scanned in shared
// memory (Hillis-Steele, two buffers), the tile offset is added on the way out.
constexpr int SC_THREADS = 1024, SC_ITEMS = 8, SC_TILE = SC_THREADS * SC_ITEMS;

__global__ void __launch_bounds__(SC_THREADS)
setup_scan_tile_sums(std::int64_t n, const unsigned long long* __restrict__ in, std::int64_t scale,
                     unsigned long long* __restrict__ tile_sum)
{
  __shared__ std::int64_t part[SC_THREADS];
  const std::int64_t base = static_cast<std::int64_t>(blockIdx.x) * SC_TILE + threadIdx.x * SC_ITEMS;
  std::int64_t sum = 0;
  for (int j = 0; j < SC_ITEMS; ++j)
    if (base + j < n)
      sum += static_cast<std::int64_t>(in[base + j]) * scale;
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int h = SC_THREADS / 2; h > 0; h >>= 1)
  {
    if (static_cast<int>(threadIdx.x) < h)
      part[threadIdx.x] += part[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    tile_sum[blockIdx.x] = static_cast<unsigned long long>(part[0]);
}

__global__ void __launch_bounds__(SC_THREADS)
setup_scan_tiles(std::int64_t n, const unsigned long long* __restrict__ in, std::int64_t scale,
                 const std::int64_t* __restrict__ tile_off, std::int64_t n_tiles, std::int64_t* __restrict__ out)
{
  __shared__ std::int64_t buf[2][SC_THREADS];
  const std::int64_t base = static_cast<std::int64_t>(blockIdx.x) * SC_TILE + threadIdx.x * SC_ITEMS;
  std::int64_t v[SC_ITEMS], sum = 0;
  for (int j = 0; j < SC_ITEMS; ++j)
  {
    v[j] = base + j < n ? static_cast<std::int64_t>(in[base + j]) * scale : 0;
    sum += v[j];
  }
  int cur = 0;
  buf[0][threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < SC_THREADS; d <<= 1) // inclusive scan of the thread sums
  {
    const std::int64_t x = buf[cur][threadIdx.x] + (static_cast<int>(threadIdx.x) >= d ? buf[cur][threadIdx.x - d] : 0);
    buf[cur ^ 1][threadIdx.x] = x;
    cur ^= 1;
    __syncthreads();
  }
  std::int64_t run = tile_off[blockIdx.x] + buf[cur][threadIdx.x] - sum;
  for (int j = 0; j < SC_ITEMS; ++j)
    if (base + j < n)
    {
      out[base + j] = run;
      run += v[j];
    }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    out[n] = tile_off[n_tiles];
}

// place pair k at the next free position of its row (cursor starts at zero)
__global__ void setup_fill(std::int64_t n_entries, const std::int32_t* __restrict__ dofmap,
                           std::int32_t n_rows, const std::int64_t* __restrict__ ptr,
                           unsigned long long* __restrict__ cursor, std::uint32_t* __restrict__ pairs)
{
  const std::int64_t k = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (k >= n_entries)
    return;
  const std::int32_t d = dofmap[k];
  if (d < n_rows)
    pairs[ptr[d] + static_cast<std::int64_t>(atomicAdd(cursor + d, 1ull))] = static_cast<std::uint32_t>(k);
}

// ascending pairs per row (insertion sort: rows hold a few dozen pairs)
__global__ void setup_sort(std::int32_t n_rows, const std::int64_t* __restrict__ ptr,
                           std::uint32_t* __restrict__ pairs)
{
  const std::int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows)
    return;
  std::uint32_t* p = pairs + ptr[r];
  const int n = static_cast<int>(ptr[r + 1] - ptr[r]);
  for (int i = 1; i < n; ++i)
  {
    const std::uint32_t v = p[i];
    int j = i - 1;
    while (j >= 0 && p[j] > v)
    {
      p[j + 1] = p[j];
      --j;
    }
    p[j + 1] = v;
  }
}

// cells per slice = longest cell list of its 32 rows
__global__ void setup_widths(std::int32_t n_rows, std::int32_t n_slices,
                             const std::int64_t* __restrict__ ptr, unsigned long long* __restrict__ wa)
{
  const std::int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slices)
    return;
  std::int64_t m = 0;
  for (std::int32_t r = 32 * s; r < min(n_rows, 32 * s + 32); ++r)
    m = max(m, ptr[r + 1] - ptr[r]);
  wa[s] = static_cast<unsigned long long>(m);
}

// rotated slot words: thread = (slice, lane). The row's columns are read from the padded SELL
// list (ascending for k < len). flags[0] is set when a vertex is missing from the row or an offset
// does not fit a byte.
__global__ void setup_adjrot(std::int32_t n_rows, std::int32_t n_slices,
                             const std::int32_t* __restrict__ dofmap,
                             const std::int64_t* __restrict__ rowptr,
                             const std::int64_t* __restrict__ mat_off,
                             const std::int32_t* __restrict__ cols_sell,
                             const std::int64_t* __restrict__ ptr,
                             const std::uint32_t* __restrict__ pairs,
                             const std::int64_t* __restrict__ adj_off,
                             std::uint32_t* __restrict__ adjrot, int* __restrict__ flags)
{
  const std::int64_t t = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  const std::int32_t s = static_cast<std::int32_t>(t >> 5);
  const int lane = static_cast<int>(t & 31);
  if (s >= n_slices)
    return;
  const std::int32_t r = 32 * s + lane;
  const bool live = r < n_rows;
  const std::int64_t ao = adj_off[s], mo = mat_off[s];
  const int wa = static_cast<int>((adj_off[s + 1] - ao) >> 5);
  const int len = live ? static_cast<int>(rowptr[r + 1] - rowptr[r]) : 0;
  const int alen = live ? static_cast<int>(ptr[r + 1] - ptr[r]) : 0;
  const std::int32_t* rc = cols_sell + mo + lane; // entry k at rc[k * 32]
  for (int k = 0; k < wa; ++k)
  {
    std::uint32_t word = ADJ_INVALID_DEV;
    if (k < alen)
    {
      const std::uint32_t pair = pairs[ptr[r] + k];
      const std::int64_t cell = pair >> 2;
      const int li = pair & 3;
      word = 0;
      for (int q = 0; q < 4; ++q)
      {
        const std::int32_t col = dofmap[cell * 4 + ((li + q) & 3)];
        int lo = 0, hi = len; // lower_bound over the row's columns
        while (lo < hi)
        {
          const int mid = (lo + hi) >> 1;
          if (rc[mid * 32] < col)
            lo = mid + 1;
          else
            hi = mid;
        }
        if (lo >= len || rc[lo * 32] != col || lo >= 255)
          flags[0] = 1;
        word |= static_cast<std::uint32_t>(lo & 0xFF) << (8 * q);
      }
    }
    adjrot[ao + static_cast<std::int64_t>(k) * 32 + lane] = word;
  }
}

// ---- column side ---------------------------------------------------------------------------------
// widths of the matrix slices: setup_widths on rowptr. Padded columns: thread = (slice, lane);
// positions past the row's length repeat its first column (0 past n_rows), as build_sell_layout does.
__global__ void setup_sell_cols(std::int32_t n_rows, std::int32_t n_slices, const std::int64_t* __restrict__ rowptr,
                                const std::int32_t* __restrict__ cols, const std::int64_t* __restrict__ mat_off,
                                std::int32_t* __restrict__ cols_sell)
{
  const std::int64_t t = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  const std::int32_t s = static_cast<std::int32_t>(t >> 5);
  const int lane = static_cast<int>(t & 31);
  if (s >= n_slices)
    return;
  const std::int32_t r = 32 * s + lane;
  const std::int64_t mo = mat_off[s], w = (mat_off[s + 1] - mo) >> 5;
  const std::int64_t len = r < n_rows ? rowptr[r + 1] - rowptr[r] : 0;
  const std::int32_t pad = len > 0 ? cols[rowptr[r]] : 0;
  for (std::int64_t k = 0; k < w; ++k)
    cols_sell[mo + k * 32 + lane] = k < len ? cols[rowptr[r] + k] : pad;
}

// compress_columns, first half: one thread per slice decides every (slice, k) and counts the
// explicit ones. A partial last slice stays explicit (row + delta may overrun).
__global__ void setup_cdelta(std::int32_t n_rows, std::int64_t n_cols, std::int32_t n_slices,
                             const std::int64_t* __restrict__ rowptr, const std::int64_t* __restrict__ mat_off,
                             const std::int32_t* __restrict__ cols_sell, std::int32_t* __restrict__ cdelta,
                             unsigned long long* __restrict__ xcnt)
{
  const std::int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slices)
    return;
  const std::int64_t mo = mat_off[s], w = (mat_off[s + 1] - mo) >> 5;
  const std::int32_t r0 = 32 * s;
  const bool full = r0 + 32 <= n_rows;
  unsigned long long nx = 0;
  for (std::int64_t k = 0; k < w; ++k)
  {
    bool uniform = full, have = false;
    std::int64_t d = 0;
    for (int lane = 0; lane < 32 && uniform; ++lane)
    {
      const std::int32_t r = r0 + lane;
      if (k >= rowptr[r + 1] - rowptr[r])
        continue; // padding: free to choose
      const std::int64_t dl = static_cast<std::int64_t>(cols_sell[mo + k * 32 + lane]) - r;
      if (!have)
        d = dl, have = true;
      else if (dl != d)
        uniform = false;
    }
    if (uniform && !have)
      d = 0; // all padding: point at the row itself
    if (uniform && (r0 + d < 0 || static_cast<std::int64_t>(r0) + 31 + d >= n_cols))
      uniform = false;
    cdelta[mo / 32 + k] = uniform ? static_cast<std::int32_t>(d) : INT_MIN;
    nx += uniform ? 0 : 1;
  }
  xcnt[s] = nx;
}

// compress_columns, second half: the explicit index lines, thread = (slice, lane)
__global__ void setup_colsx(std::int32_t n_slices, const std::int64_t* __restrict__ mat_off,
                            const std::int32_t* __restrict__ cols_sell, const std::int32_t* __restrict__ cdelta,
                            const std::int64_t* __restrict__ xoff, std::int32_t* __restrict__ colsx)
{
  const std::int64_t t = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  const std::int32_t s = static_cast<std::int32_t>(t >> 5);
  const int lane = static_cast<int>(t & 31);
  if (s >= n_slices)
    return;
  const std::int64_t mo = mat_off[s], w = (mat_off[s + 1] - mo) >> 5;
  std::int64_t j = 0;
  for (std::int64_t k = 0; k < w; ++k)
    if (cdelta[mo / 32 + k] == INT_MIN)
    {
      colsx[xoff[s] + j * 32 + lane] = cols_sell[mo + k * 32 + lane];
      ++j;
    }
}

// build_slice_order without clustering: interior[s] = 1 when no stored column of the slice is a ghost
__global__ void setup_slice_flags(std::int32_t n_rows, std::int32_t n_slices, const std::int64_t* __restrict__ mat_off,
                                  const std::int32_t* __restrict__ cols_sell, unsigned long long* __restrict__ interior)
{
  const std::int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slices)
    return;
  unsigned long long in = 1;
  for (std::int64_t q = mat_off[s]; q < mat_off[s + 1]; ++q)
    if (cols_sell[q] >= n_rows)
    {
      in = 0;
      break;
    }
  interior[s] = in;
}

// interior slices first (ascending), then the ghost-reading ones (ascending); pos = exclusive scan
// of the interior flags, pos[n_slices] = number of interior slices
__global__ void setup_slice_order(std::int32_t n_slices, const unsigned long long* __restrict__ interior,
                                  const std::int64_t* __restrict__ pos, std::int32_t* __restrict__ order)
{
  const std::int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slices)
    return;
  const std::int64_t at = interior[s] ? pos[s] : pos[n_slices] + (s - pos[s]);
  order[at] = s;
}

// Ascending union of the dofs of the row's cells, built by sorted insertion into a thread-local
// list (P1 rows hold ~15 columns, P3 vertex rows 175). Returns the count, -1 when CAP is exceeded.
constexpr int SU_MAX_COLS = 192;
__device__ __forceinline__ int row_columns(std::int32_t r, int nd, const std::int32_t* __restrict__ dofmap,
                                           const std::int64_t* __restrict__ ptr,
                                           const std::uint32_t* __restrict__ pairs, std::int32_t* list)
{
  int n = 0;
  for (std::int64_t q = ptr[r]; q < ptr[r + 1]; ++q)
  {
    const std::int64_t cell = pairs[q] / static_cast<std::uint32_t>(nd);
    for (int j = 0; j < nd; ++j)
    {
      const std::int32_t col = dofmap[cell * nd + j];
      int lo = 0, hi = n;
      while (lo < hi)
      {
        const int mid = (lo + hi) >> 1;
        if (list[mid] < col)
          lo = mid + 1;
        else
          hi = mid;
      }
      if (lo < n && list[lo] == col)
        continue;
      if (n == SU_MAX_COLS)
        return -1;
      for (int t = n; t > lo; --t)
        list[t] = list[t - 1];
      list[lo] = col;
      ++n;
    }
  }
  return n;
}

// columns per row; flags[2] is set when a row has more than SU_MAX_COLS columns
__global__ void setup_pattern_count(std::int32_t n_rows, int nd, const std::int32_t* __restrict__ dofmap,
                                    const std::int64_t* __restrict__ ptr,
                                    const std::uint32_t* __restrict__ pairs,
                                    unsigned long long* __restrict__ cnt, int* __restrict__ flags)
{
  const std::int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows)
    return;
  std::int32_t list[SU_MAX_COLS];
  const int n = row_columns(r, nd, dofmap, ptr, pairs, list);
  if (n < 0)
    flags[2] = 1;
  cnt[r] = static_cast<unsigned long long>(n < 0 ? 0 : n);
}

__global__ void setup_pattern_fill(std::int32_t n_rows, int nd, const std::int32_t* __restrict__ dofmap,
                                   const std::int64_t* __restrict__ ptr,
                                   const std::uint32_t* __restrict__ pairs,
                                   const std::int64_t* __restrict__ rowptr, std::int32_t* __restrict__ cols)
{
  const std::int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows)
    return;
  std::int32_t list[SU_MAX_COLS];
  const int n = row_columns(r, nd, dofmap, ptr, pairs, list);
  std::int32_t* out = cols + rowptr[r];
  for (int k = 0; k < n; ++k)
    out[k] = list[k];
}

// P2/P3 form of the slot map: adj = the pair itself (cell * nd + local index), adjso = the nd in-row
// offsets of the cell's dofs packed so_bits wide (layout.cpp build_sell_layout). thread = (slice, lane).
__global__ void setup_adj_pk(std::int32_t n_rows, std::int32_t n_slices, int nd, int so_bits,
                             const std::int32_t* __restrict__ dofmap, const std::int64_t* __restrict__ rowptr,
                             const std::int64_t* __restrict__ mat_off, const std::int32_t* __restrict__ cols_sell,
                             const std::int64_t* __restrict__ ptr, const std::uint32_t* __restrict__ pairs,
                             const std::int64_t* __restrict__ adj_off, std::uint32_t* __restrict__ adj,
                             std::uint32_t* __restrict__ adjso, int* __restrict__ flags)
{
  const std::int64_t t = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  const std::int32_t s = static_cast<std::int32_t>(t >> 5);
  const int lane = static_cast<int>(t & 31);
  if (s >= n_slices)
    return;
  const std::int32_t r = 32 * s + lane;
  const bool live = r < n_rows;
  const std::int64_t ao = adj_off[s], mo = mat_off[s];
  const int wa = static_cast<int>((adj_off[s + 1] - ao) >> 5);
  const int len = live ? static_cast<int>(rowptr[r + 1] - rowptr[r]) : 0;
  const int alen = live ? static_cast<int>(ptr[r + 1] - ptr[r]) : 0;
  const int per_word = 32 / so_bits, so_words = (nd + per_word - 1) / per_word;
  const std::int32_t* rc = cols_sell + mo + lane; // entry k at rc[k * 32]
  for (int k = 0; k < wa; ++k)
  {
    const bool on = k < alen;
    const std::uint32_t pair = on ? pairs[ptr[r] + k] : ADJ_INVALID_DEV;
    adj[ao + static_cast<std::int64_t>(k) * 32 + lane] = pair;
    const std::int64_t cell = pair / static_cast<std::uint32_t>(nd);
    for (int wd = 0; wd < so_words; ++wd)
    {
      std::uint32_t word = 0;
      if (on)
        for (int q = 0; q < per_word; ++q)
        {
          const int j = wd * per_word + q;
          if (j >= nd)
            break;
          const std::int32_t col = dofmap[cell * nd + j];
          int lo = 0, hi = len; // lower_bound over the row's columns
          while (lo < hi)
          {
            const int mid = (lo + hi) >> 1;
            if (rc[mid * 32] < col)
              lo = mid + 1;
            else
              hi = mid;
          }
          if (lo >= len || rc[lo * 32] != col || lo >= (1 << so_bits))
            flags[0] = 1;
          word |= static_cast<std::uint32_t>(lo) << (q * so_bits);
        }
      adjso[(ao + static_cast<std::int64_t>(k) * 32) * so_words + wd * 32 + lane] = word;
    }
  }
}

// The greedy star walk of layout.cpp build_walk: start at the row's first cell; next = the
// unvisited cell sharing most vertices with the current one, ties to the earlier cell; vertices
// that stay keep their register position, new ones take the freed positions in ascending order,
// in the cell's rotation order. flags[1] is set when a row has more than SU_MAX_CELLS cells.
__global__ void setup_walk(std::int32_t n_rows, std::int32_t n_slices,
                           const std::int64_t* __restrict__ ptr,
                           const std::int64_t* __restrict__ adj_off,
                           const std::uint32_t* __restrict__ adjrot, std::uint32_t* __restrict__ walk,
                           int* __restrict__ flags)
{
  const std::int64_t t = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  const std::int32_t s = static_cast<std::int32_t>(t >> 5);
  const int lane = static_cast<int>(t & 31);
  if (s >= n_slices)
    return;
  const std::int32_t r = 32 * s + lane;
  const std::int64_t ao = adj_off[s];
  const int wa = static_cast<int>((adj_off[s + 1] - ao) >> 5);
  const int c = r < n_rows ? static_cast<int>(ptr[r + 1] - ptr[r]) : 0;
  std::uint32_t* out = walk + ao + lane;
  for (int k = c; k < wa; ++k) // padding steps
    out[static_cast<std::int64_t>(k) * 32] = ADJ_INVALID_DEV;
  if (c == 0)
    return;
  if (c > SU_MAX_CELLS)
  {
    flags[1] = 1;
    return;
  }
  std::uint32_t o[SU_MAX_CELLS]; // the three non-owner offsets of every cell (bytes 0..2)
  for (int j = 0; j < c; ++j)
    o[j] = adjrot[ao + static_cast<std::int64_t>(j) * 32 + lane] >> 8;
  auto shared = [&](std::uint32_t a, std::uint32_t b) {
    int n = 0;
#pragma unroll
    for (int x = 0; x < 3; ++x)
#pragma unroll
      for (int y = 0; y < 3; ++y)
        n += ((a >> (8 * x)) & 0xFFu) == ((b >> (8 * y)) & 0xFFu) ? 1 : 0;
    return n;
  };
  unsigned long long visited = 1ull;
  int pos[3] = {static_cast<int>(o[0] & 0xFFu), static_cast<int>((o[0] >> 8) & 0xFFu),
                static_cast<int>((o[0] >> 16) & 0xFFu)};
  int cur = 0;
  out[0] = pos[0] | (pos[1] << 8) | (pos[2] << 16) | (7u << 24);
  for (int step = 1; step < c; ++step)
  {
    int best = -1, best_sh = -1;
    for (int j = 0; j < c; ++j)
    {
      if ((visited >> j) & 1ull)
        continue;
      const int sh = shared(o[cur], o[j]);
      if (sh > best_sh)
        best = j, best_sh = sh;
      if (sh >= 2)
        break;
    }
    bool held[3] = {false, false, false}, old[3] = {false, false, false};
    for (int q = 0; q < 3; ++q)
      for (int p = 0; p < 3; ++p)
        if (!held[p] && !old[q] && pos[p] == static_cast<int>((o[best] >> (8 * q)) & 0xFFu))
          held[p] = true, old[q] = true;
    unsigned mask = 0;
    int p = 0;
    for (int q = 0; q < 3; ++q)
    {
      if (old[q])
        continue;
      while (held[p])
        ++p;
      pos[p] = static_cast<int>((o[best] >> (8 * q)) & 0xFFu);
      held[p] = true;
      mask |= 1u << p;
    }
    out[static_cast<std::int64_t>(step) * 32] = pos[0] | (pos[1] << 8) | (pos[2] << 16) | (mask << 24);
    visited |= 1ull << best;
    cur = best;
  }
}

// ------------------------------------------------------------------------------------------
// Edge rings of layout.cpp build_rings (SellLayout::ring, the column-major elasticity kernel of
// assemble_ring.cu) on the device, one thread per row, the same chains byte for byte: for column k
// the cells that hold k as a chain of vertices; a chain starts at the lowest unvisited cell with a
// vertex no other unvisited cell of the ring shares (that vertex first), else at the lowest
// unvisited cell walking towards its lower face neighbour, and continues through the lowest
// unvisited cell that holds the current vertex. Returns the number of bytes (<= 2 * cells).
// ------------------------------------------------------------------------------------------
__device__ inline int ring_chain(const std::uint32_t* o, int c, int k, std::uint8_t* bytes)
{
  std::uint8_t va[SU_MAX_CELLS], vb[SU_MAX_CELLS];
  int n = 0;
  for (int j = 0; j < c; ++j)
    for (int t = 0; t < 3; ++t)
      if (static_cast<int>((o[j] >> (8 * t)) & 0xFFu) == k)
      {
        va[n] = static_cast<std::uint8_t>((o[j] >> (8 * ((t + 1) % 3))) & 0xFFu);
        vb[n] = static_cast<std::uint8_t>((o[j] >> (8 * ((t + 2) % 3))) & 0xFFu);
        ++n;
      }
  unsigned long long used = 0ull;
  auto degree = [&](int v) {
    int d = 0;
    for (int j = 0; j < n; ++j)
      d += !((used >> j) & 1ull) && (va[j] == v || vb[j] == v) ? 1 : 0;
    return d;
  };
  auto next_with = [&](int v, int skip) {
    for (int j = 0; j < n; ++j)
      if (!((used >> j) & 1ull) && j != skip && (va[j] == v || vb[j] == v))
        return j;
    return -1;
  };
  int left = n, nb = 0;
  while (left > 0)
  {
    int start = -1, v0 = 0, v1 = 0;
    for (int j = 0; j < n && start < 0; ++j)
    {
      if ((used >> j) & 1ull)
        continue;
      if (degree(va[j]) == 1)
        start = j, v0 = va[j], v1 = vb[j];
      else if (degree(vb[j]) == 1)
        start = j, v0 = vb[j], v1 = va[j];
    }
    if (start < 0)
    {
      for (int j = 0; j < n && start < 0; ++j)
        if (!((used >> j) & 1ull))
          start = j;
      const int ja = next_with(va[start], start), jb = next_with(vb[start], start);
      if (ja <= jb)
        v1 = va[start], v0 = vb[start];
      else
        v1 = vb[start], v0 = va[start];
    }
    bytes[nb++] = static_cast<std::uint8_t>(v0 | 0x80);
    bytes[nb++] = static_cast<std::uint8_t>(v1);
    used |= 1ull << start, --left;
    int cur = v1;
    for (int j = next_with(cur, -1); j >= 0; j = next_with(cur, -1))
    {
      cur = va[j] == cur ? vb[j] : va[j];
      bytes[nb++] = static_cast<std::uint8_t>(cur);
      used |= 1ull << j, --left;
    }
  }
  return nb;
}

// PASS 0: ns32[mat_off[s]/32 + k] = longest chain of column k over the rows of slice s (atomicMax).
// PASS 1: the chain bytes, four to a word, into ring at ring_off[s] + (words before column k + q)*32 + lane;
// every word of the slice is written (padding 0x80). flags[1] is set when a row has more than
// SU_MAX_CELLS cells or a chain of more than 255 bytes.
template <int PASS>
__global__ void setup_rings(std::int32_t n_rows, std::int32_t n_slices, const std::int64_t* __restrict__ ptr,
                            const std::int64_t* __restrict__ rowptr, const std::int64_t* __restrict__ mat_off,
                            const std::int64_t* __restrict__ adj_off, const std::uint32_t* __restrict__ adjrot,
                            int* __restrict__ ns32, const std::uint8_t* __restrict__ ring_ns,
                            const std::int64_t* __restrict__ ring_off, std::uint32_t* __restrict__ ring,
                            int* __restrict__ flags)
{
  const std::int64_t t = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  const std::int32_t s = static_cast<std::int32_t>(t >> 5);
  const int lane = static_cast<int>(t & 31);
  if (s >= n_slices)
    return;
  const std::int32_t r = 32 * s + lane;
  const bool live = r < n_rows;
  const std::int64_t ao = adj_off[s];
  int c = live ? static_cast<int>(ptr[r + 1] - ptr[r]) : 0;
  const int len = live ? static_cast<int>(rowptr[r + 1] - rowptr[r]) : 0;
  if (c > SU_MAX_CELLS)
  {
    flags[1] = 1;
    c = 0;
  }
  std::uint32_t o[SU_MAX_CELLS]; // the three non-owner offsets of every cell (bytes 0..2)
  for (int j = 0; j < c; ++j)
    o[j] = adjrot[ao + static_cast<std::int64_t>(j) * 32 + lane] >> 8;
  const std::int64_t k0 = mat_off[s] >> 5;
  const int w = static_cast<int>((mat_off[s + 1] - mat_off[s]) >> 5);
  std::uint8_t bytes[2 * SU_MAX_CELLS];
  if constexpr (PASS == 0)
  {
    for (int k = 0; k < len; ++k)
    {
      const int nb = ring_chain(o, c, k, bytes);
      if (nb > 255)
        flags[1] = 1;
      else if (nb > 0)
        atomicMax(ns32 + k0 + k, nb);
    }
  }
  else
  {
    std::int64_t base = ring_off[s] + lane;
    for (int k = 0; k < w; ++k)
    {
      const int ns = ring_ns[k0 + k];
      const int nw = (ns + 3) >> 2;
      const int nb = (k < len && ns > 0) ? ring_chain(o, c, k, bytes) : 0;
      for (int q = 0; q < nw; ++q)
      {
        std::uint32_t word = 0;
        for (int u = 0; u < 4; ++u)
          word |= static_cast<std::uint32_t>(4 * q + u < nb ? bytes[4 * q + u] : 0x80u) << (8 * u);
        ring[base + static_cast<std::int64_t>(q) * 32] = word;
      }
      base += static_cast<std::int64_t>(nw) * 32;
    }
  }
}

// ring_ns (uint8) from the pass-0 maxima and the ring words per lane of every slice.
__global__ void setup_ring_words(std::int32_t n_slices, const std::int64_t* __restrict__ mat_off,
                                 const int* __restrict__ ns32, std::uint8_t* __restrict__ ring_ns,
                                 unsigned long long* __restrict__ words)
{
  const std::int64_t s = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (s >= n_slices)
    return;
  const std::int64_t k0 = mat_off[s] >> 5;
  const int w = static_cast<int>((mat_off[s + 1] - mat_off[s]) >> 5);
  unsigned long long n = 0;
  for (int k = 0; k < w; ++k)
  {
    const int ns = ns32[k0 + k];
    ring_ns[k0 + k] = static_cast<std::uint8_t>(ns);
    n += static_cast<unsigned long long>((ns + 3) >> 2);
  }
  words[s] = n;
}

} // namespace

#ifndef PTB_HOST_EMU // host side: device build only
// Builds adj_off, adjrot and (if want_walk) walk on the device from the dofmap and the already
// uploaded column side (rowptr, mat_off, padded columns). Returns false when the device build does
// not apply (a row with too many cells / offsets beyond a byte): the caller then uses the host build.
namespace
{
// out[0..n] = exclusive prefix sum of in[0..n) * scale (out[n] = total): one CTA for short inputs,
// the three tile passes above otherwise.
void device_scan(ptb_ctx* c, std::int64_t n, const unsigned long long* in, std::int64_t* out, std::int64_t scale)
{
  if (n <= SC_TILE)
  {
    setup_scan<<<1, 1024, 0, c->stream>>>(n, in, out, scale);
    c->launches += 1;
    return;
  }
  const std::int64_t n_tiles = (n + SC_TILE - 1) / SC_TILE;
  DevBuf<unsigned long long> tile_sum;
  DevBuf<std::int64_t> tile_off;
  tile_sum.alloc(static_cast<std::size_t>(n_tiles));
  tile_off.alloc(static_cast<std::size_t>(n_tiles) + 1);
  setup_scan_tile_sums<<<static_cast<unsigned>(n_tiles), SC_THREADS, 0, c->stream>>>(n, in, scale, tile_sum.p);
  setup_scan<<<1, 1024, 0, c->stream>>>(n_tiles, tile_sum.p, tile_off.p, 1);
  setup_scan_tiles<<<static_cast<unsigned>(n_tiles), SC_THREADS, 0, c->stream>>>(n, in, scale, tile_off.p, n_tiles, out);
  PTB_CUDA(cudaGetLastError());
  PTB_CUDA(cudaStreamSynchronize(c->stream)); // the tile buffers die here
  c->launches += 3;
}

// dof -> (cell, local index) pairs of the owned rows, ascending per row: ptr [N + 1], pairs [ptr[N]].
void build_pairs(ptb_ctx* c, DevBuf<std::int64_t>& ptr, DevBuf<std::uint32_t>& pairs)
{
  const std::int32_t N = c->n_owned;
  const std::int64_t n_entries = c->n_cells * c->nd;
  DevBuf<unsigned long long> cnt;
  cnt.alloc(static_cast<std::size_t>(N));
  cnt.zero(c->stream);
  ptr.alloc(static_cast<std::size_t>(N) + 1);
  const int ge = static_cast<int>((n_entries + SU_THREADS - 1) / SU_THREADS);
  setup_count<<<ge, SU_THREADS, 0, c->stream>>>(n_entries, c->dofmap.p, N, cnt.p);
  device_scan(c, N, cnt.p, ptr.p, 1);
  std::int64_t n_pairs = 0;
  PTB_CUDA(cudaMemcpyAsync(&n_pairs, ptr.p + N, sizeof(n_pairs), cudaMemcpyDeviceToHost, c->stream));
  PTB_CUDA(cudaStreamSynchronize(c->stream));
  pairs.alloc(static_cast<std::size_t>(n_pairs));
  cnt.zero(c->stream); // reused as the fill cursor
  setup_fill<<<ge, SU_THREADS, 0, c->stream>>>(n_entries, c->dofmap.p, N, ptr.p, cnt.p, pairs.p);
  setup_sort<<<(N + SU_THREADS - 1) / SU_THREADS, SU_THREADS, 0, c->stream>>>(N, ptr.p, pairs.p);
  PTB_CUDA(cudaGetLastError());
  PTB_CUDA(cudaStreamSynchronize(c->stream)); // cnt dies here
  c->launches += 3;
}
} // namespace

// The sparsity pattern of the owned rows built on the device and downloaded (any order). Returns
// false when a row has more than SU_MAX_COLS columns (the caller builds the pattern on the host).
bool gpu_build_pattern(ptb_ctx* c, std::vector<std::int64_t>& rowptr, std::vector<std::int32_t>& cols,
                       DevBuf<std::int64_t>& rp, DevBuf<std::int32_t>& cl)
{
  const std::int32_t N = c->n_owned;
  DevBuf<std::int64_t> ptr;
  DevBuf<std::uint32_t> pairs;
  DevBuf<unsigned long long> cnt;
  DevBuf<int> flags;
  build_pairs(c, ptr, pairs);
  cnt.alloc(static_cast<std::size_t>(N));
  rp.alloc(static_cast<std::size_t>(N) + 1);
  flags.alloc(3);
  flags.zero(c->stream);
  const int gr = (N + SU_THREADS - 1) / SU_THREADS;
  setup_pattern_count<<<gr, SU_THREADS, 0, c->stream>>>(N, c->nd, c->dofmap.p, ptr.p, pairs.p, cnt.p, flags.p);
  device_scan(c, N, cnt.p, rp.p, 1);
  rowptr.resize(static_cast<std::size_t>(N) + 1);
  int h_flags[3] = {0, 0, 0};
  PTB_CUDA(cudaMemcpyAsync(rowptr.data(), rp.p, rowptr.size() * sizeof(std::int64_t), cudaMemcpyDeviceToHost,
                           c->stream));
  PTB_CUDA(cudaMemcpyAsync(h_flags, flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, c->stream));
  PTB_CUDA(cudaStreamSynchronize(c->stream));
  c->launches += 1;
  if (h_flags[2] != 0)
    return false;
  cl.alloc(static_cast<std::size_t>(rowptr[N]));
  setup_pattern_fill<<<gr, SU_THREADS, 0, c->stream>>>(N, c->nd, c->dofmap.p, ptr.p, pairs.p, rp.p, cl.p);
  PTB_CUDA(cudaGetLastError());
  cols.resize(static_cast<std::size_t>(rowptr[N]));
  PTB_CUDA(cudaMemcpyAsync(cols.data(), cl.p, cols.size() * sizeof(std::int32_t), cudaMemcpyDeviceToHost,
                           c->stream));
  PTB_CUDA(cudaStreamSynchronize(c->stream));
  c->launches += 1;
  return true;
}

// The column side of ptb_set_pattern from a CSR pattern that is already on the device: c->rowptr,
// mat_off, cols (SELL-32, padded), for scalar problems cdelta / colsx / xoff, and the slice order.
// Sets n_slices, max_w, cols_explicit_frac, n_interior_slices.
void gpu_setup_columns(ptb_ctx* c, DevBuf<std::int64_t>& rp, const DevBuf<std::int32_t>& cl,
                       std::vector<std::int64_t>& h_mat_off)
{
  const std::int32_t N = c->n_owned, S = (N + 31) / 32;
  const std::int64_t n_cols = static_cast<std::int64_t>(N) + c->n_ghost;
  c->n_slices = S;
  DevBuf<unsigned long long> w;
  w.alloc(static_cast<std::size_t>(S));
  const int gs = (S + SU_THREADS - 1) / SU_THREADS;
  const int gl = static_cast<int>((static_cast<std::int64_t>(S) * 32 + SU_THREADS - 1) / SU_THREADS);
  setup_widths<<<gs, SU_THREADS, 0, c->stream>>>(N, S, rp.p, w.p);
  c->mat_off.alloc(static_cast<std::size_t>(S) + 1);
  device_scan(c, S, w.p, c->mat_off.p, 32);
  std::int64_t n_sell = 0;
  std::vector<unsigned long long> h_w(static_cast<std::size_t>(S));
  PTB_CUDA(cudaMemcpyAsync(&n_sell, c->mat_off.p + S, sizeof(n_sell), cudaMemcpyDeviceToHost, c->stream));
  PTB_CUDA(cudaMemcpyAsync(h_w.data(), w.p, h_w.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                           c->stream));
  PTB_CUDA(cudaStreamSynchronize(c->stream));
  c->max_w = 0;
  h_mat_off.assign(static_cast<std::size_t>(S) + 1, 0); // host copy of the offsets (O(n_slices))
  for (std::int32_t s = 0; s < S; ++s)
  {
    c->max_w = std::max(c->max_w, static_cast<int>(h_w[s]));
    h_mat_off[s + 1] = h_mat_off[s] + 32 * static_cast<std::int64_t>(h_w[s]);
  }
  c->cols.alloc(static_cast<std::size_t>(n_sell));
  setup_sell_cols<<<gl, SU_THREADS, 0, c->stream>>>(N, S, rp.p, cl.p, c->mat_off.p, c->cols.p);
  c->launches += 2;
  if (c->bs == 1)
  {
    c->cdelta.alloc(static_cast<std::size_t>(n_sell / 32));
    setup_cdelta<<<gs, SU_THREADS, 0, c->stream>>>(N, n_cols, S, rp.p, c->mat_off.p, c->cols.p, c->cdelta.p, w.p);
    c->xoff.alloc(static_cast<std::size_t>(S) + 1);
    device_scan(c, S, w.p, c->xoff.p, 32);
    std::int64_t n_x = 0;
    PTB_CUDA(cudaMemcpyAsync(&n_x, c->xoff.p + S, sizeof(n_x), cudaMemcpyDeviceToHost, c->stream));
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    c->colsx.alloc(static_cast<std::size_t>(n_x));
    if (n_x > 0)
      setup_colsx<<<gl, SU_THREADS, 0, c->stream>>>(S, c->mat_off.p, c->cols.p, c->cdelta.p, c->xoff.p, c->colsx.p);
    c->cols_explicit_frac = n_sell > 0 ? static_cast<double>(n_x) / static_cast<double>(n_sell) : 0.0;
    c->launches += 2;
  }
  DevBuf<std::int64_t> pos;
  pos.alloc(static_cast<std::size_t>(S) + 1);
  c->slice_order.alloc(static_cast<std::size_t>(S));
  setup_slice_flags<<<gs, SU_THREADS, 0, c->stream>>>(N, S, c->mat_off.p, c->cols.p, w.p);
  device_scan(c, S, w.p, pos.p, 1);
  setup_slice_order<<<gs, SU_THREADS, 0, c->stream>>>(S, w.p, pos.p, c->slice_order.p);
  PTB_CUDA(cudaGetLastError());
  std::int64_t n_int = 0;
  PTB_CUDA(cudaMemcpyAsync(&n_int, pos.p + S, sizeof(n_int), cudaMemcpyDeviceToHost, c->stream));
  PTB_CUDA(cudaStreamSynchronize(c->stream)); // w, pos die here
  c->n_interior_slices = static_cast<std::int32_t>(n_int);
  c->launches += 2;
  // the kernels read rowptr from the context: take the device copy over instead of re-uploading it
  std::swap(c->rowptr.p, rp.p);
  std::swap(c->rowptr.n, rp.n);
}

// P2/P3: adj_off, adj and adjso on the device (so_bits = 8: rows of at most 256 columns). Returns
// false when the pattern does not cover a cell (the host build then reports it).
bool gpu_setup_pk(ptb_ctx* c, int* max_wa)
{
  const std::int32_t N = c->n_owned, S = c->n_slices;
  const int so_bits = 8, so_words = (c->nd + 3) / 4;
  DevBuf<unsigned long long> wa;
  DevBuf<std::int64_t> ptr;
  DevBuf<std::uint32_t> pairs;
  DevBuf<int> flags;
  build_pairs(c, ptr, pairs);
  flags.alloc(2);
  flags.zero(c->stream);
  wa.alloc(static_cast<std::size_t>(S));
  setup_widths<<<(S + SU_THREADS - 1) / SU_THREADS, SU_THREADS, 0, c->stream>>>(N, S, ptr.p, wa.p);
  c->adj_off.alloc(static_cast<std::size_t>(S) + 1);
  device_scan(c, S, wa.p, c->adj_off.p, 32);
  std::int64_t n_adj = 0;
  PTB_CUDA(cudaMemcpyAsync(&n_adj, c->adj_off.p + S, sizeof(n_adj), cudaMemcpyDeviceToHost, c->stream));
  std::vector<unsigned long long> h_wa(static_cast<std::size_t>(S));
  PTB_CUDA(cudaMemcpyAsync(h_wa.data(), wa.p, h_wa.size() * sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, c->stream));
  PTB_CUDA(cudaStreamSynchronize(c->stream));
  *max_wa = 0;
  for (unsigned long long v : h_wa)
    *max_wa = std::max(*max_wa, static_cast<int>(v));
  c->adj.alloc(static_cast<std::size_t>(n_adj));
  c->adjso.alloc(static_cast<std::size_t>(n_adj) * so_words);
  const int gl = static_cast<int>((static_cast<std::int64_t>(S) * 32 + SU_THREADS - 1) / SU_THREADS);
  setup_adj_pk<<<gl, SU_THREADS, 0, c->stream>>>(N, S, c->nd, so_bits, c->dofmap.p, c->rowptr.p, c->mat_off.p,
                                                 c->cols.p, ptr.p, pairs.p, c->adj_off.p, c->adj.p, c->adjso.p,
                                                 flags.p);
  PTB_CUDA(cudaGetLastError());
  int h_flags[2] = {0, 0};
  PTB_CUDA(cudaMemcpyAsync(h_flags, flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, c->stream));
  PTB_CUDA(cudaStreamSynchronize(c->stream));
  c->launches += 2;
  return h_flags[0] == 0;
}

bool gpu_setup_p1(ptb_ctx* c, bool want_walk, bool want_rings, int* max_wa)
{
  const std::int32_t N = c->n_owned, S = c->n_slices;
  DevBuf<unsigned long long> wa;
  DevBuf<std::int64_t> ptr;
  DevBuf<std::uint32_t> pairs;
  DevBuf<int> flags;
  build_pairs(c, ptr, pairs);
  flags.alloc(2);
  flags.zero(c->stream);
  wa.alloc(static_cast<std::size_t>(S));
  setup_widths<<<(S + SU_THREADS - 1) / SU_THREADS, SU_THREADS, 0, c->stream>>>(N, S, ptr.p, wa.p);
  c->adj_off.alloc(static_cast<std::size_t>(S) + 1);
  device_scan(c, S, wa.p, c->adj_off.p, 32);
  std::int64_t n_adj = 0;
  PTB_CUDA(cudaMemcpyAsync(&n_adj, c->adj_off.p + S, sizeof(n_adj), cudaMemcpyDeviceToHost, c->stream));
  std::vector<unsigned long long> h_wa(static_cast<std::size_t>(S));
  PTB_CUDA(cudaMemcpyAsync(h_wa.data(), wa.p, h_wa.size() * sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, c->stream));
  PTB_CUDA(cudaStreamSynchronize(c->stream));
  *max_wa = 0;
  for (unsigned long long v : h_wa)
    *max_wa = std::max(*max_wa, static_cast<int>(v));
  c->adjrot.alloc(static_cast<std::size_t>(n_adj));
  const int gl = static_cast<int>((static_cast<std::int64_t>(S) * 32 + SU_THREADS - 1) / SU_THREADS);
  setup_adjrot<<<gl, SU_THREADS, 0, c->stream>>>(N, S, c->dofmap.p, c->rowptr.p, c->mat_off.p, c->cols.p,
                                                 ptr.p, pairs.p, c->adj_off.p, c->adjrot.p, flags.p);
  if (want_walk)
  {
    c->walk.alloc(static_cast<std::size_t>(n_adj));
    setup_walk<<<gl, SU_THREADS, 0, c->stream>>>(N, S, ptr.p, c->adj_off.p, c->adjrot.p, c->walk.p, flags.p);
  }
  c->ring.release(), c->ring_off.release(), c->ring_ns.release();
  c->ring_max_words = 0;
  if (want_rings && c->max_w <= 127)
  {
    // edge rings of the column-major elasticity kernel (assemble_ring.cu), as layout.cpp build_rings
    const std::size_t n_cols = c->cols.n / 32;
    DevBuf<int> ns32;
    DevBuf<unsigned long long> words;
    ns32.alloc(n_cols);
    ns32.zero(c->stream);
    words.alloc(static_cast<std::size_t>(S));
    c->ring_ns.alloc(n_cols);
    c->ring_off.alloc(static_cast<std::size_t>(S) + 1);
    setup_rings<0><<<gl, SU_THREADS, 0, c->stream>>>(N, S, ptr.p, c->rowptr.p, c->mat_off.p, c->adj_off.p,
                                                     c->adjrot.p, ns32.p, nullptr, nullptr, nullptr, flags.p);
    setup_ring_words<<<(S + SU_THREADS - 1) / SU_THREADS, SU_THREADS, 0, c->stream>>>(S, c->mat_off.p, ns32.p,
                                                                                     c->ring_ns.p, words.p);
    device_scan(c, S, words.p, c->ring_off.p, 32);
    std::int64_t n_ring = 0;
    std::vector<unsigned long long> h_words(static_cast<std::size_t>(S));
    PTB_CUDA(cudaMemcpyAsync(&n_ring, c->ring_off.p + S, sizeof(n_ring), cudaMemcpyDeviceToHost, c->stream));
    PTB_CUDA(cudaMemcpyAsync(h_words.data(), words.p, h_words.size() * sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, c->stream));
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    for (unsigned long long v : h_words)
      c->ring_max_words = std::max(c->ring_max_words, static_cast<int>(v));
    c->ring.alloc(static_cast<std::size_t>(n_ring));
    setup_rings<1><<<gl, SU_THREADS, 0, c->stream>>>(N, S, ptr.p, c->rowptr.p, c->mat_off.p, c->adj_off.p,
                                                     c->adjrot.p, nullptr, c->ring_ns.p, c->ring_off.p, c->ring.p,
                                                     flags.p);
    c->ring_bytes_per_row = N ? 4.0 * static_cast<double>(n_ring) / N : 0.0; // padded words, not chain bytes
    c->launches += 3;
  }
  PTB_CUDA(cudaGetLastError());
  int h_flags[2] = {0, 0};
  PTB_CUDA(cudaMemcpyAsync(h_flags, flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, c->stream));
  PTB_CUDA(cudaStreamSynchronize(c->stream));
  c->launches += want_walk ? 3 : 2;
  if (h_flags[0] != 0 || h_flags[1] != 0)
  {
    c->ring.release(), c->ring_off.release(), c->ring_ns.release();
    return false;
  }
  return true;
}
#endif // PTB_HOST_EMU

} // namespace ptb
