// TEST INFRASTRUCTURE -- runs the source of the ascending-cell-order P1 kernels (csrc/assemble.cu:
// the default elasticity matrix kernel, the default vector kernels, the facet kernel) on the host,
// like emu_kernels.cpp does for the walk kernels. Nothing here is linked into the product
// libraries; the product path never sees PTB_HOST_EMU.
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
static inline void __syncwarp(unsigned = 0xffffffffu) {} // per-lane private shared memory in these kernels
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline double __drcp_rn(double d) { return 1.0 / d; }
// declared for kernels of the file that this harness does not run (matrix-free action)
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }

namespace ptb
{
namespace
{
alignas(16) double smem[232448 / 8];
}
} // namespace ptb

#include "../../performance-test_b200/csrc/assemble.cu"

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

int emu_p1_matrix(int bs, int32_t n_rows, int32_t n_slices, int max_w, const uint8_t* bc,
                  const int64_t* rowptr, const int64_t* mat_off, const int64_t* adj_off,
                  const int32_t* cols, const uint32_t* adjrot, const double* xdof, double* vals,
                  double* dinv)
{
  using namespace ptb;
  MatrixArgs A{};
  A.n_rows = n_rows, A.n_slices = n_slices, A.bc = bc, A.rowptr = rowptr, A.mat_off = mat_off;
  A.adj_off = adj_off, A.cols = cols, A.adjrot = adjrot, A.xdof = xdof, A.max_w = max_w;
  A.vals = vals, A.dinv = dinv;
  if (bs == 1)
    emu_launch(assemble_matrix_p1<1>, (n_slices + 3) / 4, MAT_THREADS_1, A);
  else
    emu_launch(assemble_matrix_p1<3>, (n_slices + 1) / 2, MAT_THREADS_3, A);
  return 0;
}

int emu_p1_vector(int bs, int32_t n_rows, int32_t n_slices, int max_w, const uint8_t* bc,
                  const int64_t* mat_off, const int64_t* adj_off, const int32_t* cols,
                  const uint32_t* adjrot, const double* xdof, const double* f, double* b)
{
  using namespace ptb;
  VectorArgs A{};
  A.n_rows = n_rows, A.n_slices = n_slices, A.bc = bc, A.adj_off = adj_off, A.adjrot = adjrot;
  A.xdof = xdof, A.mat_off = mat_off, A.cols = cols, A.max_w = max_w, A.f = f, A.b = b;
  if (bs == 1)
    emu_launch(assemble_vector_p1<1>, (n_slices + 3) / 4, MAT_THREADS_1, A);
  else
    emu_launch(assemble_vector_p1<3>, (n_slices + 1) / 2, MAT_THREADS_3, A);
  return 0;
}

// exterior facets of the scalar P1 space (assemble_facets_p1): adds g v ds into b
int emu_p1_facets(int32_t n_frows, const double* xyz4, const int32_t* x_dofmap,
                  const int32_t* dofmap, const uint8_t* bc, const int32_t* frow_ids,
                  const int32_t* frow_ptr, const int32_t* fent, const double* g, double* b)
{
  using namespace ptb;
  FacetArgs F{n_frows, xyz4, x_dofmap, dofmap, bc, frow_ids, frow_ptr, fent, g, b};
  if (n_frows > 0)
    emu_launch(assemble_facets_p1, (n_frows + 127) / 128, 128, F);
  return 0;
}
}
