// TEST INFRASTRUCTURE -- runs the *source* of the P1 walk kernels (performance-test_b200/csrc/
// assemble_walk.cu, assemble_gwalk.cu) on the host, thread by thread, so that their indexing,
// step decoding and shared-memory layouts can be checked without a GPU. Nothing here is linked
// into the product libraries; the product path never sees PTB_HOST_EMU.
//
// Execution model: one CTA at a time; its threads are std::threads that share the "shared
// memory" array and meet at __syncthreads() (std::barrier). Kernels without barriers would also
// run sequentially, but one model for all keeps the shim small. Warp-level primitives are not
// provided: the kernels compiled here do not use any.
#define PTB_HOST_EMU 1
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include <cuda_runtime.h> // vector types, host-side spelling of __global__ & co. (attributes g++ ignores)

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
template <typename T>
static inline T __ldg(const T* p)
{
  return *p;
}
// warp shuffle (only the matrix-free action uses one): the 32 threads of a warp meet at a barrier
static std::barrier<>* emu_warp[32];
static double emu_xd[32][32];
static inline double __shfl_xor_sync(unsigned, double v, int o)
{
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  emu_xd[w][l] = v;
  emu_warp[w]->arrive_and_wait();
  const double r = emu_xd[w][l ^ o];
  emu_warp[w]->arrive_and_wait();
  return r;
}

static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp[threadIdx.x >> 5]->arrive_and_wait(); }

namespace ptb
{
namespace
{
alignas(16) double smem[232448 / 8]; // extern __shared__ double smem[] of the kernels (227 KB)
}
} // namespace ptb

#include "../../performance-test_b200/csrc/assemble_walk.cu"
#include "../../performance-test_b200/csrc/assemble_gwalk.cu"
#include "../../performance-test_b200/csrc/assemble_ring.cu"

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
    std::vector<std::unique_ptr<std::barrier<>>> wb;
    for (unsigned w = 0; w < (block + 31) / 32; ++w)
    {
      wb.push_back(std::make_unique<std::barrier<>>(32));
      emu_warp[w] = wb.back().get();
    }
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t)
      th.emplace_back([=] {
        threadIdx.x = t, blockIdx.x = b;
        kernel(args...);
        emu_barrier->arrive_and_drop(); // a thread that has returned no longer takes part
      });
    for (auto& x : th)
      x.join();
  }
}
} // namespace

extern "C" {

// variant: 0 walk<1,true>  1 walk<2,false>  2 walk3<true>  6 walk<1,true,EXACT>
int emu_assemble_matrix(int variant, int32_t n_rows, int32_t n_slices, int max_w, int bs,
                        const uint8_t* bc, const int64_t* rowptr, const int64_t* mat_off,
                        const int64_t* adj_off, const int32_t* cols, const double* xdof,
                        const uint32_t* walk, const uint32_t* walk1, const int64_t* walk1_off,
                        double* vals, double* dinv)
{
  using namespace ptb;
  MatrixArgs A{};
  A.n_rows = n_rows, A.n_slices = n_slices, A.bc = bc, A.rowptr = rowptr, A.mat_off = mat_off;
  A.adj_off = adj_off, A.cols = cols, A.xdof = xdof, A.max_w = max_w, A.vals = vals, A.dinv = dinv;
  switch (variant)
  {
  case 0: emu_launch(assemble_matrix_p1_walk<1, true>, n_slices, 32, A, walk); break;
  case 6: emu_launch(assemble_matrix_p1_walk<1, true, true>, n_slices, 32, A, walk); break; // cross_rn
  case 1: emu_launch(assemble_matrix_p1_walk<2, false>, (n_slices + 1) / 2, 64, A, walk); break;
  case 2:
    if (bs != 3)
      return 2;
    emu_launch(assemble_matrix_p1_walk3<true>, n_slices, 96, A, walk);
    break;
  default: return 1;
  }
  return 0;
}

// assemble_matrix_p1_ring3<warps> (elasticity, column-major along the edge rings)
int emu_assemble_matrix_ring(int warps, int32_t n_rows, int32_t n_slices, int max_w, const uint8_t* bc,
                             const int64_t* rowptr, const int64_t* mat_off, const int32_t* cols,
                             const double* xdof, const uint32_t* ring, const int64_t* ring_off,
                             const uint8_t* ring_ns, int max_rw, double* vals, double* dinv)
{
  using namespace ptb;
  MatrixArgs A{};
  A.n_rows = n_rows, A.n_slices = n_slices, A.bc = bc, A.rowptr = rowptr, A.mat_off = mat_off;
  A.cols = cols, A.xdof = xdof, A.max_w = max_w, A.vals = vals, A.dinv = dinv;
  if (warps == 1)
    emu_launch(assemble_matrix_p1_ring3<1, false>, n_slices, 32, A, ring, ring_off, ring_ns, max_rw);
  else if (warps == 4)
    emu_launch(assemble_matrix_p1_ring3<4, false>, (n_slices + 3) / 4, 128, A, ring, ring_off, ring_ns, max_rw);
  else
    return 1;
  return 0;
}

int emu_assemble_vector(int bs, int warps, int32_t n_rows, int32_t n_slices, int max_w,
                        const uint8_t* bc, const int64_t* mat_off, const int32_t* cols,
                        const double* xdof, const double* f, const uint32_t* walk1,
                        const int64_t* walk1_off, double* b)
{
  using namespace ptb;
  VectorArgs A{};
  A.n_rows = n_rows, A.n_slices = n_slices, A.bc = bc, A.mat_off = mat_off, A.cols = cols;
  A.xdof = xdof, A.max_w = max_w, A.f = f, A.b = b;
  const unsigned nw = static_cast<unsigned>(n_slices); // one warp per slice, all components
  (void)bs;
  if (bs == 1 && warps == 4)
    emu_launch(assemble_vector_p1_gwalk<1, 4>, (nw + 3) / 4, 128, A, walk1, walk1_off);
  else if (bs == 1 && warps == 1)
    emu_launch(assemble_vector_p1_gwalk<1, 1>, nw, 32, A, walk1, walk1_off);
  else if (bs == 3 && warps == 4)
    emu_launch(assemble_vector_p1_gwalk<3, 4>, (nw + 3) / 4, 128, A, walk1, walk1_off);
  else if (bs == 3 && warps == 8)
    emu_launch(assemble_vector_p1_gwalk<3, 8>, (nw + 7) / 8, 256, A, walk1, walk1_off);
  else
    return 1;
  return 0;
}
}
