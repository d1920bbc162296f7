// TEST INFRASTRUCTURE -- runs the source of the device-side problem-data kernels
// (csrc/problem_data.cu) on the host. Nothing here is linked into the product libraries; the product
// path never sees PTB_HOST_EMU.
#define PTB_HOST_EMU 1
#include <cmath>
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

struct EmuIdx
{
  unsigned x = 0, y = 0, z = 0;
};
static EmuIdx threadIdx, blockIdx, blockDim, gridDim;
// the non-contracting intrinsics: this file is built without FMA contraction (-ffp-contract=off)
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
using std::exp;
using std::fabs;
using std::sin;
using std::sqrt;

#include "../../performance-test_b200/csrc/problem_data.cu"

namespace
{
// the kernels have no barriers and no shared memory: their threads run one after the other
template <typename K, typename... Args>
void emu_launch(K kernel, std::int64_t n_threads, unsigned block, Args... args)
{
  blockDim.x = block;
  gridDim.x = static_cast<unsigned>((n_threads + block - 1) / block);
  for (unsigned b = 0; b < gridDim.x; ++b)
    for (unsigned t = 0; t < block; ++t)
    {
      blockIdx.x = b, threadIdx.x = t;
      kernel(args...);
    }
}
} // namespace

extern "C" {

int emu_locate_bc(int64_t n_cells, int problem, int order, int nd, const double* xyz4, const int32_t* x_dofmap,
                  const int32_t* dofmap, uint8_t* bc)
{
  using namespace ptb;
  emu_launch(locate_bc_facets, n_cells * 4, PD_THREADS, n_cells, problem, order, nd, xyz4, x_dofmap, dofmap, bc);
  return 0;
}

int emu_interpolate_source(int64_t n, int problem, const double* X, int stride, double* f, double* g)
{
  using namespace ptb;
  emu_launch(interpolate_source, n, PD_THREADS, n, problem, X, stride, f, g);
  return 0;
}
}
