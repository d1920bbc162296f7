// TEST INFRASTRUCTURE -- runs the source of the device-side mesh / P1 dofmap generator (csrc/box.cu)
// on the host. Nothing here is linked into the product libraries; the product path never sees
// PTB_HOST_EMU.
#define PTB_HOST_EMU 1
#include <climits>
#include <cmath>
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

struct EmuIdx
{
  unsigned x = 0, y = 0, z = 0;
};
static EmuIdx threadIdx, blockIdx, blockDim, gridDim;
static inline double __dmul_rn(double a, double b) { return a * b; }

#include "../../performance-test_b200/csrc/box.cu"

namespace
{
// no barriers, no shared memory: the threads run one after the other
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

// dims = nx, ny, nz, l0, l1, G0, G1, Glow, Ghigh (BoxDims as gpu_create_box_p1 fills it)
int emu_create_box_p1(const int64_t* dims, double* xyz3, double* xyz4, int32_t* dof_vertex, int32_t* x_dofmap,
                      int32_t* dofmap)
{
  using namespace ptb;
  BoxDims B{dims[0], dims[1], dims[2], dims[3], dims[4], dims[5], dims[6], dims[7], dims[8]};
  const std::int64_t nvp = (B.nx + 1) * (B.ny + 1);
  const double hx = 1.0 / static_cast<double>(B.nx), hy = 1.0 / static_cast<double>(B.ny),
               hz = 1.0 / static_cast<double>(B.nz);
  emu_launch(box_vertices, nvp * (B.l1 - B.l0 + 1), BX_THREADS, B, hx, hy, hz, xyz3, xyz4, dof_vertex);
  emu_launch(box_cells_p1, B.nx * B.ny * (B.l1 - B.l0), BX_THREADS, B, x_dofmap, dofmap);
  return 0;
}

int emu_gather_dofmap_rows(int64_t n, int nd, const int32_t* cells, const int32_t* dofmap, int32_t* out)
{
  using namespace ptb;
  emu_launch(gather_dofmap_rows, n * nd, BX_THREADS, n, nd, cells, dofmap, out);
  return 0;
}
}
