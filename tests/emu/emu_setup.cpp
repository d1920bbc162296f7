// TEST INFRASTRUCTURE -- runs the source of the device-side P1 setup kernels (csrc/setup.cu) on the
// host. Nothing here is linked into the product libraries; the product path never sees PTB_HOST_EMU.
#define PTB_HOST_EMU 1
#include <algorithm>
#include <barrier>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

struct EmuIdx
{
  unsigned x = 0, y = 0, z = 0;
};
static thread_local EmuIdx threadIdx, blockIdx;
static EmuIdx blockDim, gridDim;
static std::barrier<>* emu_barrier = nullptr;
static inline void __syncthreads() { emu_barrier->arrive_and_wait(); }
// the threads of a CTA run concurrently here, so the atomics are real ones
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v)
{
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline int atomicMax(int* p, int v)
{
  int old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
  {
  }
  return old;
}
using std::max;
using std::min;

// the fixture passes a copy of csrc/setup.cu in which `__shared__` reads `static` (one CTA at a time)
#include PTB_EMU_SETUP_SOURCE

namespace
{
template <typename K, typename... Args>
void emu_launch(K kernel, unsigned grid, unsigned block, Args... args)
{
  gridDim.x = grid, blockDim.x = block;
  for (unsigned b = 0; b < grid; ++b)
  {
    std::barrier<> bar(block);
    emu_barrier = &bar;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < block; ++t)
      th.emplace_back([=] {
        threadIdx.x = t, blockIdx.x = b;
        kernel(args...);
        emu_barrier->arrive_and_drop();
      });
    for (auto& x : th)
      x.join();
  }
}
} // namespace

namespace
{
// device_scan of setup.cu: one CTA for short inputs, the three tile passes otherwise
void emu_scan(std::int64_t n, const unsigned long long* in, std::int64_t* out, std::int64_t scale)
{
  using namespace ptb;
  if (n <= SC_TILE)
  {
    emu_launch(setup_scan, 1, 1024, n, in, out, scale);
    return;
  }
  const std::int64_t n_tiles = (n + SC_TILE - 1) / SC_TILE;
  std::vector<unsigned long long> tile_sum(static_cast<std::size_t>(n_tiles), 0);
  std::vector<std::int64_t> tile_off(static_cast<std::size_t>(n_tiles) + 1, -1);
  emu_launch(setup_scan_tile_sums, static_cast<unsigned>(n_tiles), SC_THREADS, n, in, scale, tile_sum.data());
  emu_launch(setup_scan, 1, 1024, n_tiles, (const unsigned long long*)tile_sum.data(), tile_off.data(),
             static_cast<std::int64_t>(1));
  emu_launch(setup_scan_tiles, static_cast<unsigned>(n_tiles), SC_THREADS, n, in, scale,
             (const std::int64_t*)tile_off.data(), n_tiles, out);
}

// dof -> (cell, local index) pairs as build_pairs (setup.cu) launches them
void emu_pairs(int64_t n_entries, const int32_t* dofmap, int32_t n_rows, int shuffle, std::vector<std::int64_t>& ptr,
               std::vector<std::uint32_t>& pairs)
{
  using namespace ptb;
  const unsigned ge = static_cast<unsigned>((n_entries + SU_THREADS - 1) / SU_THREADS);
  std::vector<unsigned long long> cnt(static_cast<std::size_t>(n_rows), 0);
  ptr.assign(static_cast<std::size_t>(n_rows) + 1, -1);
  emu_launch(setup_count, ge, SU_THREADS, n_entries, dofmap, n_rows, cnt.data());
  emu_scan(static_cast<std::int64_t>(n_rows), (const unsigned long long*)cnt.data(), ptr.data(), static_cast<std::int64_t>(1));
  pairs.assign(static_cast<std::size_t>(ptr[n_rows]), 0xFFFFFFFFu);
  std::fill(cnt.begin(), cnt.end(), 0ull);
  emu_launch(setup_fill, ge, SU_THREADS, n_entries, dofmap, n_rows, (const std::int64_t*)ptr.data(), cnt.data(),
             pairs.data());
  if (shuffle)
    for (std::int32_t r = 0; r < n_rows; ++r)
      std::reverse(pairs.begin() + ptr[r], pairs.begin() + ptr[r + 1]);
  emu_launch(setup_sort, (n_rows + SU_THREADS - 1) / SU_THREADS, SU_THREADS, n_rows, (const std::int64_t*)ptr.data(),
             pairs.data());
}
} // namespace

extern "C" {

// the prefix sum alone (both routes of device_scan)
int emu_scan_only(int64_t n, const unsigned long long* in, int64_t scale, int64_t* out)
{
  emu_scan(n, in, out, scale);
  return 0;
}

// The launch sequence of gpu_build_pattern (setup.cu). rowptr [n_rows + 1] is always written; cols
// (capacity cap) when the total fits (returns -1 otherwise). flags[2] as in setup.cu.
int emu_build_pattern(int64_t n_cells, int nd, const int32_t* dofmap, int32_t n_rows, int64_t cap, int64_t* rowptr,
                      int32_t* cols, int* flags)
{
  using namespace ptb;
  std::vector<std::int64_t> ptr;
  std::vector<std::uint32_t> pairs;
  emu_pairs(n_cells * nd, dofmap, n_rows, 0, ptr, pairs);
  std::vector<unsigned long long> cnt(static_cast<std::size_t>(n_rows), 0);
  flags[0] = flags[1] = flags[2] = 0;
  const unsigned gr = (n_rows + SU_THREADS - 1) / SU_THREADS;
  emu_launch(setup_pattern_count, gr, SU_THREADS, n_rows, nd, dofmap, (const std::int64_t*)ptr.data(),
             (const std::uint32_t*)pairs.data(), cnt.data(), flags);
  emu_scan(static_cast<std::int64_t>(n_rows), (const unsigned long long*)cnt.data(), rowptr, static_cast<std::int64_t>(1));
  if (flags[2] != 0 || rowptr[n_rows] > cap)
    return -1;
  emu_launch(setup_pattern_fill, gr, SU_THREADS, n_rows, nd, dofmap, (const std::int64_t*)ptr.data(),
             (const std::uint32_t*)pairs.data(), (const std::int64_t*)rowptr, cols);
  return 0;
}

// The launch sequence of gpu_setup_columns (setup.cu) for a scalar problem. mat_off, xoff, order
// [n_slices (+ 1)] are always written; cols_sell / cdelta (capacity cap, cap / 32) and colsx
// (capacity capx) when the totals fit (returns -1 otherwise). *n_interior = interior slices.
int emu_setup_columns(int32_t n_rows, int64_t n_cols, int32_t n_slices, const int64_t* rowptr, const int32_t* cols,
                      int64_t cap, int64_t capx, int64_t* mat_off, int32_t* cols_sell, int32_t* cdelta, int64_t* xoff,
                      int32_t* colsx, int32_t* order, int32_t* n_interior)
{
  using namespace ptb;
  std::vector<unsigned long long> w(static_cast<std::size_t>(n_slices), 0);
  const unsigned gs = (n_slices + SU_THREADS - 1) / SU_THREADS;
  const unsigned gl = static_cast<unsigned>((static_cast<std::int64_t>(n_slices) * 32 + SU_THREADS - 1) / SU_THREADS);
  emu_launch(setup_widths, gs, SU_THREADS, n_rows, n_slices, rowptr, w.data());
  emu_scan(static_cast<std::int64_t>(n_slices), (const unsigned long long*)w.data(), mat_off, static_cast<std::int64_t>(32));
  if (mat_off[n_slices] > cap)
    return -1;
  emu_launch(setup_sell_cols, gl, SU_THREADS, n_rows, n_slices, rowptr, cols, (const std::int64_t*)mat_off, cols_sell);
  emu_launch(setup_cdelta, gs, SU_THREADS, n_rows, n_cols, n_slices, rowptr, (const std::int64_t*)mat_off,
             (const std::int32_t*)cols_sell, cdelta, w.data());
  emu_scan(static_cast<std::int64_t>(n_slices), (const unsigned long long*)w.data(), xoff, static_cast<std::int64_t>(32));
  if (xoff[n_slices] > capx)
    return -1;
  emu_launch(setup_colsx, gl, SU_THREADS, n_slices, (const std::int64_t*)mat_off, (const std::int32_t*)cols_sell,
             (const std::int32_t*)cdelta, (const std::int64_t*)xoff, colsx);
  std::vector<std::int64_t> pos(static_cast<std::size_t>(n_slices) + 1, -1);
  emu_launch(setup_slice_flags, gs, SU_THREADS, n_rows, n_slices, (const std::int64_t*)mat_off,
             (const std::int32_t*)cols_sell, w.data());
  emu_scan(static_cast<std::int64_t>(n_slices), (const unsigned long long*)w.data(), pos.data(), static_cast<std::int64_t>(1));
  emu_launch(setup_slice_order, gs, SU_THREADS, n_slices, (const unsigned long long*)w.data(),
             (const std::int64_t*)pos.data(), order);
  *n_interior = static_cast<std::int32_t>(pos[n_slices]);
  return 0;
}

// The launch sequence of gpu_setup_pk (setup.cu): adj_off always; adj (capacity cap) and adjso
// (capacity cap * so_words) when the total fits (returns -1 otherwise).
int emu_setup_pk(int64_t n_cells, int nd, const int32_t* dofmap, int32_t n_rows, int32_t n_slices,
                 const int64_t* rowptr, const int64_t* mat_off, const int32_t* cols_sell, int64_t cap,
                 int64_t* adj_off, uint32_t* adj, uint32_t* adjso, int* flags)
{
  using namespace ptb;
  std::vector<unsigned long long> wa(static_cast<std::size_t>(n_slices), 0);
  std::vector<std::int64_t> ptr;
  std::vector<std::uint32_t> pairs;
  emu_pairs(n_cells * nd, dofmap, n_rows, 1, ptr, pairs);
  emu_launch(setup_widths, (n_slices + SU_THREADS - 1) / SU_THREADS, SU_THREADS, n_rows, n_slices,
             (const std::int64_t*)ptr.data(), wa.data());
  emu_scan(static_cast<std::int64_t>(n_slices), (const unsigned long long*)wa.data(), adj_off, static_cast<std::int64_t>(32));
  if (adj_off[n_slices] > cap)
    return -1;
  flags[0] = flags[1] = 0;
  const unsigned gl = static_cast<unsigned>((static_cast<std::int64_t>(n_slices) * 32 + SU_THREADS - 1) / SU_THREADS);
  emu_launch(setup_adj_pk, gl, SU_THREADS, n_rows, n_slices, nd, 8, dofmap, rowptr, mat_off, cols_sell,
             (const std::int64_t*)ptr.data(), (const std::uint32_t*)pairs.data(), (const std::int64_t*)adj_off, adj,
             adjso, flags);
  return 0;
}

// The launch sequence of gpu_setup_p1 (setup.cu) with host vectors in place of the device buffers.
// adj_off [n_slices + 1] is always written; adjrot / walk hold `cap` words each and are written
// when the device-side total fits (returns -1 otherwise). flags[0..1] as in setup.cu.
// shuffle != 0 reverses the pairs of every row before the per-row sort: the order in which the
// atomics of setup_fill land is arbitrary on a GPU, the result must not depend on it.
int emu_setup_p1(int64_t n_cells, const int32_t* dofmap, int32_t n_rows, int32_t n_slices,
                 const int64_t* rowptr, const int64_t* mat_off, const int32_t* cols_sell, int shuffle,
                 int64_t cap, int64_t* adj_off, uint32_t* adjrot, uint32_t* walk, int* flags)
{
  using namespace ptb;
  std::vector<unsigned long long> wa(static_cast<std::size_t>(n_slices), 0);
  std::vector<std::int64_t> ptr;
  std::vector<std::uint32_t> pairs;
  emu_pairs(n_cells * 4, dofmap, n_rows, shuffle, ptr, pairs);
  emu_launch(setup_widths, (n_slices + SU_THREADS - 1) / SU_THREADS, SU_THREADS, n_rows, n_slices,
             (const std::int64_t*)ptr.data(), wa.data());
  emu_scan(static_cast<std::int64_t>(n_slices), (const unsigned long long*)wa.data(), adj_off, static_cast<std::int64_t>(32));
  if (adj_off[n_slices] > cap)
    return -1;
  flags[0] = flags[1] = 0;
  const unsigned gl = static_cast<unsigned>((static_cast<std::int64_t>(n_slices) * 32 + SU_THREADS - 1) / SU_THREADS);
  emu_launch(setup_adjrot, gl, SU_THREADS, n_rows, n_slices, dofmap, rowptr, mat_off, cols_sell,
             (const std::int64_t*)ptr.data(), (const std::uint32_t*)pairs.data(), (const std::int64_t*)adj_off, adjrot,
             flags);
  emu_launch(setup_walk, gl, SU_THREADS, n_rows, n_slices, (const std::int64_t*)ptr.data(),
             (const std::int64_t*)adj_off, (const std::uint32_t*)adjrot, walk, flags);
  return 0;
}

// The ring part of gpu_setup_p1 (setup.cu): pairs -> adj_off / adjrot -> setup_rings<0> -> setup_ring_words
// -> scan -> setup_rings<1>. ring_off [n_slices + 1] and ring_ns [mat_off[n_slices] / 32] are always
// written; ring (capacity cap words) when the total fits (returns -1 otherwise). flags[0..1] as in setup.cu.
int emu_setup_p1_rings(int64_t n_cells, const int32_t* dofmap, int32_t n_rows, int32_t n_slices,
                       const int64_t* rowptr, const int64_t* mat_off, const int32_t* cols_sell, int shuffle,
                       int64_t cap, int64_t* ring_off, uint8_t* ring_ns, uint32_t* ring, int* flags)
{
  using namespace ptb;
  std::vector<unsigned long long> wa(static_cast<std::size_t>(n_slices), 0);
  std::vector<std::int64_t> ptr, adj_off(static_cast<std::size_t>(n_slices) + 1, -1);
  std::vector<std::uint32_t> pairs;
  emu_pairs(n_cells * 4, dofmap, n_rows, shuffle, ptr, pairs);
  emu_launch(setup_widths, (n_slices + SU_THREADS - 1) / SU_THREADS, SU_THREADS, n_rows, n_slices,
             (const std::int64_t*)ptr.data(), wa.data());
  emu_scan(static_cast<std::int64_t>(n_slices), (const unsigned long long*)wa.data(), adj_off.data(), static_cast<std::int64_t>(32));
  std::vector<std::uint32_t> adjrot(static_cast<std::size_t>(adj_off[n_slices]), 0xDEADBEEFu);
  flags[0] = flags[1] = 0;
  const unsigned gl = static_cast<unsigned>((static_cast<std::int64_t>(n_slices) * 32 + SU_THREADS - 1) / SU_THREADS);
  emu_launch(setup_adjrot, gl, SU_THREADS, n_rows, n_slices, dofmap, rowptr, mat_off, cols_sell,
             (const std::int64_t*)ptr.data(), (const std::uint32_t*)pairs.data(), (const std::int64_t*)adj_off.data(),
             adjrot.data(), flags);
  const std::size_t n_cols = static_cast<std::size_t>(mat_off[n_slices] / 32);
  std::vector<int> ns32(n_cols, 0);
  std::vector<unsigned long long> words(static_cast<std::size_t>(n_slices), 0);
  emu_launch(setup_rings<0>, gl, SU_THREADS, n_rows, n_slices, (const std::int64_t*)ptr.data(), rowptr, mat_off,
             (const std::int64_t*)adj_off.data(), (const std::uint32_t*)adjrot.data(), ns32.data(),
             (const std::uint8_t*)nullptr, (const std::int64_t*)nullptr, (std::uint32_t*)nullptr, flags);
  emu_launch(setup_ring_words, (n_slices + SU_THREADS - 1) / SU_THREADS, SU_THREADS, n_slices, mat_off,
             (const int*)ns32.data(), ring_ns, words.data());
  emu_scan(static_cast<std::int64_t>(n_slices), (const unsigned long long*)words.data(), ring_off, static_cast<std::int64_t>(32));
  if (ring_off[n_slices] > cap)
    return -1;
  emu_launch(setup_rings<1>, gl, SU_THREADS, n_rows, n_slices, (const std::int64_t*)ptr.data(), rowptr, mat_off,
             (const std::int64_t*)adj_off.data(), (const std::uint32_t*)adjrot.data(), (int*)nullptr,
             (const std::uint8_t*)ring_ns, (const std::int64_t*)ring_off, ring, flags);
  return 0;
}
}
