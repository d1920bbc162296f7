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
//   pattern count / scan / fill  the sparsity pattern itself (common/intmaps.cpp build_pattern; the
//                                reference builds it in fem::create_matrix, poisson_problem.cpp:122-123,
//                                inside the timed assembly stage): per owned row the ascending union of
//                                the dofs of its cells, any Lagrange order (ptb_build_pattern)
// The column side (mat_off, padded columns, column compression, slice order) stays on the host: it
// only needs the caller's CSR pattern, no adjacency.
// NOT YET RUN ON A GPU (written after the round's GPU budget was spent).
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

} // namespace

#ifndef PTB_HOST_EMU // host side: device build only
// Builds adj_off, adjrot and (if want_walk) walk on the device from the dofmap and the already
// uploaded column side (rowptr, mat_off, padded columns). Returns false when the device build does
// not apply (a row with too many cells / offsets beyond a byte): the caller then uses the host build.
namespace
{
// dof -> (cell, local index) pairs of the owned rows, ascending per row: ptr [N + 1], pairs [ptr[N]].
// Five launches.
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
  setup_scan<<<1, 1024, 0, c->stream>>>(N, cnt.p, ptr.p, 1);
  std::int64_t n_pairs = 0;
  PTB_CUDA(cudaMemcpyAsync(&n_pairs, ptr.p + N, sizeof(n_pairs), cudaMemcpyDeviceToHost, c->stream));
  PTB_CUDA(cudaStreamSynchronize(c->stream));
  pairs.alloc(static_cast<std::size_t>(n_pairs));
  cnt.zero(c->stream); // reused as the fill cursor
  setup_fill<<<ge, SU_THREADS, 0, c->stream>>>(n_entries, c->dofmap.p, N, ptr.p, cnt.p, pairs.p);
  setup_sort<<<(N + SU_THREADS - 1) / SU_THREADS, SU_THREADS, 0, c->stream>>>(N, ptr.p, pairs.p);
  PTB_CUDA(cudaGetLastError());
  PTB_CUDA(cudaStreamSynchronize(c->stream)); // cnt dies here
  c->launches += 5;
}
} // namespace

// The sparsity pattern of the owned rows built on the device and downloaded (any order). Returns
// false when a row has more than SU_MAX_COLS columns (the caller builds the pattern on the host).
bool gpu_build_pattern(ptb_ctx* c, std::vector<std::int64_t>& rowptr, std::vector<std::int32_t>& cols)
{
  const std::int32_t N = c->n_owned;
  DevBuf<std::int64_t> ptr, rp;
  DevBuf<std::uint32_t> pairs;
  DevBuf<unsigned long long> cnt;
  DevBuf<std::int32_t> cl;
  DevBuf<int> flags;
  build_pairs(c, ptr, pairs);
  cnt.alloc(static_cast<std::size_t>(N));
  rp.alloc(static_cast<std::size_t>(N) + 1);
  flags.alloc(3);
  flags.zero(c->stream);
  const int gr = (N + SU_THREADS - 1) / SU_THREADS;
  setup_pattern_count<<<gr, SU_THREADS, 0, c->stream>>>(N, c->nd, c->dofmap.p, ptr.p, pairs.p, cnt.p, flags.p);
  setup_scan<<<1, 1024, 0, c->stream>>>(N, cnt.p, rp.p, 1);
  rowptr.resize(static_cast<std::size_t>(N) + 1);
  int h_flags[3] = {0, 0, 0};
  PTB_CUDA(cudaMemcpyAsync(rowptr.data(), rp.p, rowptr.size() * sizeof(std::int64_t), cudaMemcpyDeviceToHost,
                           c->stream));
  PTB_CUDA(cudaMemcpyAsync(h_flags, flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, c->stream));
  PTB_CUDA(cudaStreamSynchronize(c->stream));
  c->launches += 2;
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

bool gpu_setup_p1(ptb_ctx* c, bool want_walk, int* max_wa)
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
  setup_scan<<<1, 1024, 0, c->stream>>>(S, wa.p, c->adj_off.p, 32);
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
  PTB_CUDA(cudaGetLastError());
  int h_flags[2] = {0, 0};
  PTB_CUDA(cudaMemcpyAsync(h_flags, flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, c->stream));
  PTB_CUDA(cudaStreamSynchronize(c->stream));
  c->launches += want_walk ? 4 : 3;
  return h_flags[0] == 0 && h_flags[1] == 0;
}
#endif // PTB_HOST_EMU

} // namespace ptb
