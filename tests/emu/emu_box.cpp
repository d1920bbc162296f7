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
static inline double __dadd_rn(double a, double b) { return a + b; }

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

// The slab of `rank` of `nranks` exactly as gpu_create_box sets it up: make_space_dims and
// make_box_dims are the product's own host functions. dims_out receives the BoxDims (nx, ny, nz, l0,
// l1, G0, G1, Glow, Ghigh) for the test's independent check. dofmap [n_cells * nd]; dof_x
// [(n_owned + n_ghost) * 3].
int emu_create_box(int64_t nx, int64_t ny, int64_t nz, int rank, int nranks, int order, int64_t* dims_out,
                   double* xyz3, double* xyz4, int32_t* dof_vertex, int32_t* x_dofmap, int32_t* dofmap,
                   double* dof_x, int* flags)
{
  using namespace ptb;
  const SpaceDims N = make_space_dims(nx, ny, nz, order);
  const BoxDims B = make_box_dims(nx, ny, nz, rank, nranks, N);
  const std::int64_t d[9] = {B.nx, B.ny, B.nz, B.l0, B.l1, B.G0, B.G1, B.Glow, B.Ghigh};
  for (int i = 0; i < 9; ++i)
    dims_out[i] = d[i];
  const std::int64_t nvp = (B.nx + 1) * (B.ny + 1), n_cubes = B.nx * B.ny * (B.l1 - B.l0);
  const double hx = 1.0 / static_cast<double>(B.nx), hy = 1.0 / static_cast<double>(B.ny),
               hz = 1.0 / static_cast<double>(B.nz);
  emu_launch(box_vertices, nvp * (B.l1 - B.l0 + 1), BX_THREADS, B, N.PS + N.LS, hx, hy, hz, xyz3, xyz4, dof_vertex);
  flags[0] = 0;
  if (order == 1)
    emu_launch(box_cells_p1, n_cubes, BX_THREADS, B, x_dofmap, dofmap);
  else
  {
    std::vector<std::int32_t> scratch(static_cast<std::size_t>(n_cubes) * 24);
    emu_launch(box_cells_p1, n_cubes, BX_THREADS, B, x_dofmap, scratch.data());
    emu_launch(box_cells_dofmap, 6 * n_cubes, BX_THREADS, B, N, dofmap, flags);
  }
  // the coordinate kernel runs for every order here (the product launches it for order > 1 only)
  emu_launch(box_dof_coordinates, B.Ghigh - B.Glow, BX_THREADS, B, N, hx, hy, hz, dof_x);
  return 0;
}

int emu_gather_dofmap_rows(int64_t n, int nd, const int32_t* cells, const int32_t* dofmap, int32_t* out)
{
  using namespace ptb;
  emu_launch(gather_dofmap_rows, n * nd, BX_THREADS, n, nd, cells, dofmap, out);
  return 0;
}
}
