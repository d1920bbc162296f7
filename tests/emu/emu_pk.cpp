// TEST INFRASTRUCTURE -- runs the source of the P2/P3 matrix assembly kernels (csrc/assemble_pk.cu)
// on the host, like emu_kernels.cpp does for the P1 walk kernels. Nothing here is linked into the
// product libraries; the product path never sees PTB_HOST_EMU.
#define PTB_HOST_EMU 1
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
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
static inline void __syncwarp(unsigned = 0xffffffffu) {} // the accumulators are private per lane
template <typename T>
static inline T __ldg(const T* p) { return *p; }
// declared for the other kernels of the file (matrix-free action), which this harness does not run
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }

namespace ptb
{
namespace
{
alignas(16) double smem[232448 / 8];
}
} // namespace ptb

// The test fixture writes a copy of csrc/assemble_pk.cu in which `extern __shared__` (dynamic shared
// memory -> the array above) reads `extern` and every other `__shared__` (static shared memory ->
// one instance per CTA; CTAs run one at a time here) reads `static`, and passes its path.
#include PTB_EMU_PK_SOURCE

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

extern "C" {

// binned = 0: assemble_matrix_pk over all slices; 1: assemble_matrix_pk_binned, bin after bin;
// 2: cell_geometry_pk over the n_cells cells, then assemble_matrix_pk_binned<.., CELLG> bin after bin
int emu_assemble_matrix_pk(int binned, int nd, int so_bits, int32_t n_rows, int32_t n_slices,
                           int max_w, const double* xyz4, const int32_t* x_dofmap,
                           const uint8_t* bc, const int64_t* rowptr, const int64_t* mat_off,
                           const int64_t* adj_off, const int32_t* cols, const uint32_t* adj,
                           const uint32_t* adjso, int n_bins, const int32_t* bin_off,
                           const int* bin_w, const int32_t* bin_slices, double* vals, double* dinv,
                           int64_t n_cells)
{
  using namespace ptb;
  MatrixArgs A{};
  A.n_rows = n_rows, A.n_slices = n_slices, A.xyz = xyz4, A.x_dofmap = x_dofmap, A.bc = bc;
  A.rowptr = rowptr, A.mat_off = mat_off, A.adj_off = adj_off, A.cols = cols, A.adj = adj;
  A.adjso = adjso, A.max_w = max_w, A.vals = vals, A.dinv = dinv;
  const double* S = nd == 10 ? tables::S_P2 : tables::S_P3;
  auto run = [&](auto ND, auto WIDE) {
    constexpr int N = decltype(ND)::value;
    constexpr bool W = decltype(WIDE)::value;
    std::vector<double> cell_g;
    if (binned == 2)
    {
      cell_g.assign(static_cast<std::size_t>(n_cells) * 8, std::nan(""));
      emu_launch(cell_geometry_pk, static_cast<unsigned>((n_cells + 255) / 256), 256, n_cells, x_dofmap, xyz4,
                 cell_g.data());
      A.cell_g = cell_g.data();
    }
    if (!binned)
      emu_launch(assemble_matrix_pk<N, W>, (n_slices + 3) / 4, PK_THREADS, A, S);
    else if (binned == 2)
      for (int b = 0; b < n_bins; ++b)
      {
        const std::int32_t n = bin_off[b + 1] - bin_off[b];
        if (n > 0)
          emu_launch(assemble_matrix_pk_binned<N, W, true>, (n + 3) / 4, PK_THREADS, A, S,
                     bin_slices + bin_off[b], n, bin_w[b]);
      }
    else
      for (int b = 0; b < n_bins; ++b)
      {
        const std::int32_t n = bin_off[b + 1] - bin_off[b];
        if (n > 0)
          emu_launch(assemble_matrix_pk_binned<N, W>, (n + 3) / 4, PK_THREADS, A, S,
                     bin_slices + bin_off[b], n, bin_w[b]);
      }
  };
  if (nd == 10 && so_bits == 8)
    run(std::integral_constant<int, 10>{}, std::false_type{});
  else if (nd == 10)
    run(std::integral_constant<int, 10>{}, std::true_type{});
  else if (nd == 20 && so_bits == 8)
    run(std::integral_constant<int, 20>{}, std::false_type{});
  else if (nd == 20)
    run(std::integral_constant<int, 20>{}, std::true_type{});
  else
    return 1;
  return 0;
}

// cell vector + exterior facets of the P2/P3 space (assemble_vector_pk, assemble_facets_pk)
int emu_assemble_vector_pk(int nd, int32_t n_rows, int32_t n_slices, const double* xyz4,
                           const int32_t* x_dofmap, const int32_t* dofmap, const uint8_t* bc,
                           const int64_t* adj_off, const uint32_t* adj, const double* f,
                           int32_t n_frows, const int32_t* frow_ids, const int32_t* frow_ptr,
                           const int32_t* fent, const double* g, double* b)
{
  using namespace ptb;
  VectorArgs A{};
  A.n_rows = n_rows, A.n_slices = n_slices, A.xyz = xyz4, A.x_dofmap = x_dofmap, A.dofmap = dofmap;
  A.bc = bc, A.adj_off = adj_off, A.adj = adj, A.f = f, A.b = b;
  FacetArgs F{n_frows, xyz4, x_dofmap, dofmap, bc, frow_ids, frow_ptr, fent, g, b};
  if (nd == 10)
  {
    emu_launch(assemble_vector_pk<10>, (n_slices + 3) / 4, PK_THREADS, A, tables::M_P2);
    if (n_frows > 0)
      emu_launch(assemble_facets_pk<10>, (n_frows + 127) / 128, 128, F, tables::MF_P2);
  }
  else if (nd == 20)
  {
    emu_launch(assemble_vector_pk<20>, (n_slices + 3) / 4, PK_THREADS, A, tables::M_P3);
    if (n_frows > 0)
      emu_launch(assemble_facets_pk<20>, (n_frows + 127) / 128, 128, F, tables::MF_P3);
  }
  else
    return 1;
  return 0;
}
}
