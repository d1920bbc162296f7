// Device-side mesh and P1 function space of the reference's unit cube -- SURVEY 8f row 4: the
// "ZZZ Create Mesh" / "ZZZ FunctionSpace" regions (mesh.cpp:184-186 mesh::create_box(tetrahedron),
// poisson_problem.cpp:33-47 create_functionspace) generated on the device instead of being passed
// in through ptb_set_mesh / ptb_set_space, so that a C5-scale slab (198 M cells per GPU) needs no
// multi-GB host staging. Same arrays as the host stand-in (host/box_mesh.cpp, host/fem.cpp), bit
// for bit: Kuhn 6-tet split of every cube (all six share the body diagonal 0-7), vertices
// lexicographic (plane, iy, ix), x = ix * (1 / nx); z-slab partition with one ghost layer of cells
// below the slab; P1 dofs = vertices, owned planes first, then the ghost plane below, then the
// ghost plane above.
// NOT YET RUN ON A GPU (written after the round's GPU budget was spent); tests/emu runs these
// sources on the host against the stand-in's arrays.
#include "kernels.h"

namespace ptb
{
namespace
{

constexpr int BX_THREADS = 256;

struct BoxDims
{
  std::int64_t nx, ny, nz;         // global cube counts
  std::int64_t l0, l1;             // local cube layers [l0, l1); local vertex planes l0 .. l1
  std::int64_t G0, G1, Glow, Ghigh; // global vertex numbers: owned [G0, G1), ghosts [Glow, G0) and [G1, Ghigh)
};

__device__ __forceinline__ std::int32_t box_to_local(const BoxDims& B, std::int64_t g)
{
  if (g >= B.G0 && g < B.G1)
    return static_cast<std::int32_t>(g - B.G0);
  if (g >= B.Glow && g < B.G0)
    return static_cast<std::int32_t>((B.G1 - B.G0) + (g - B.Glow));
  if (g >= B.G1 && g < B.Ghigh)
    return static_cast<std::int32_t>((B.G1 - B.G0) + (B.G0 - B.Glow) + (g - B.G1));
  return -1;
}

// thread per local vertex: coordinates in the caller's layout [v][3] and padded [v][4], and the
// inverse of the P1 dof numbering (dof -> vertex)
__global__ void box_vertices(BoxDims B, double hx, double hy, double hz, double* __restrict__ xyz3,
                             double* __restrict__ xyz4, std::int32_t* __restrict__ dof_vertex)
{
  const std::int64_t nvx = B.nx + 1, nvp = nvx * (B.ny + 1);
  const std::int64_t v = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (v >= nvp * (B.l1 - B.l0 + 1))
    return;
  const std::int64_t ix = v % nvx, iy = (v / nvx) % (B.ny + 1), pz = v / nvp + B.l0;
  const double x = __dmul_rn(hx, static_cast<double>(ix)), y = __dmul_rn(hy, static_cast<double>(iy)),
               z = __dmul_rn(hz, static_cast<double>(pz));
  xyz3[3 * v + 0] = x, xyz3[3 * v + 1] = y, xyz3[3 * v + 2] = z;
  xyz4[4 * v + 0] = x, xyz4[4 * v + 1] = y, xyz4[4 * v + 2] = z, xyz4[4 * v + 3] = 0.0;
  const std::int32_t d = box_to_local(B, v + B.l0 * nvp);
  if (d >= 0)
    dof_vertex[d] = static_cast<std::int32_t>(v);
}

// thread per local cube: its six tetrahedra as vertex indices (x_dofmap) and as P1 dofs (dofmap)
__global__ void box_cells_p1(BoxDims B, std::int32_t* __restrict__ x_dofmap, std::int32_t* __restrict__ dofmap)
{
  // corner c of a cube has offset (c & 1, (c >> 1) & 1, (c >> 2) & 1)
  constexpr int kuhn[6][4] = {{0, 1, 3, 7}, {0, 1, 7, 5}, {0, 5, 7, 4}, {0, 3, 2, 7}, {0, 6, 4, 7}, {0, 2, 6, 7}};
  const std::int64_t nvx = B.nx + 1, nvp = nvx * (B.ny + 1);
  const std::int64_t cube = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (cube >= B.nx * B.ny * (B.l1 - B.l0))
    return;
  const std::int64_t ix = cube % B.nx, iy = (cube / B.nx) % B.ny, iz = cube / (B.nx * B.ny);
  const std::int64_t v0 = iz * nvp + iy * nvx + ix;
#pragma unroll
  for (int t = 0; t < 6; ++t)
#pragma unroll
    for (int a = 0; a < 4; ++a)
    {
      const int o = kuhn[t][a];
      const std::int64_t v = v0 + (o & 1) + ((o >> 1) & 1) * nvx + ((o >> 2) & 1) * nvp;
      x_dofmap[24 * cube + 4 * t + a] = static_cast<std::int32_t>(v);
      dofmap[24 * cube + 4 * t + a] = box_to_local(B, v + B.l0 * nvp);
    }
}

// the dofmap rows of a list of cells (thread per entry): what the host needs of a device-generated
// dofmap to build the exterior-facet row lists
__global__ void gather_dofmap_rows(std::int64_t n, int nd, const std::int32_t* __restrict__ cells,
                                   const std::int32_t* __restrict__ dofmap, std::int32_t* __restrict__ out)
{
  const std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n * nd)
    return;
  out[i] = dofmap[static_cast<std::int64_t>(cells[i / nd]) * nd + i % nd];
}

} // namespace

#ifndef PTB_HOST_EMU // host side: device build only
void launch_gather_dofmap_rows(ptb_ctx* c, std::int64_t n, const std::int32_t* cells, std::int32_t* out)
{
  const std::int64_t t = n * c->nd;
  gather_dofmap_rows<<<static_cast<unsigned>((t + BX_THREADS - 1) / BX_THREADS), BX_THREADS, 0, c->stream>>>(
      n, c->nd, cells, c->dofmap.p, out);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

// Generates the local slab of rank `rank` of `nranks` on the device and fills what ptb_set_mesh and
// the mesh-dependent half of ptb_set_space fill: n_vertices, n_cells, x_dofmap, xyz3, xyz, n_owned,
// n_ghost, dofmap, dof_vertex. Two launches.
void gpu_create_box_p1(ptb_ctx* c, std::int64_t nx, std::int64_t ny, std::int64_t nz, int rank, int nranks)
{
  // z-slabs of cube layers, as evenly as possible (host/box_mesh.cpp slab_range)
  const std::int64_t base = nz / nranks, rem = nz % nranks;
  const std::int64_t L0 = rank * base + (rank < rem ? rank : rem), L1 = L0 + base + (rank < rem ? 1 : 0);
  const bool last = rank == nranks - 1;
  const std::int64_t nvp = (nx + 1) * (ny + 1);
  BoxDims B{};
  B.nx = nx, B.ny = ny, B.nz = nz;
  B.l0 = rank > 0 ? L0 - 1 : L0, B.l1 = L1;
  B.G0 = L0 * nvp;
  B.G1 = last ? (nz + 1) * nvp : L1 * nvp;
  B.Glow = B.l0 * nvp;
  B.Ghigh = last ? B.G1 : B.G1 + nvp;
  const std::int64_t n_vertices = nvp * (B.l1 - B.l0 + 1), n_cubes = nx * ny * (B.l1 - B.l0);
  if (n_vertices > INT32_MAX || n_cubes * 24 > static_cast<std::int64_t>(UINT32_MAX))
    throw std::runtime_error("ptb_create_box_p1: local slab exceeds 32-bit local indexing");
  c->n_vertices = n_vertices, c->n_cells = 6 * n_cubes;
  c->n_owned = static_cast<std::int32_t>(B.G1 - B.G0);
  c->n_ghost = static_cast<std::int32_t>((B.G0 - B.Glow) + (B.Ghigh - B.G1));
  c->xyz3.alloc(static_cast<std::size_t>(n_vertices) * 3);
  c->xyz.alloc(static_cast<std::size_t>(n_vertices) * 4);
  c->x_dofmap.alloc(static_cast<std::size_t>(n_cubes) * 24);
  c->dofmap.alloc(static_cast<std::size_t>(n_cubes) * 24);
  c->dof_vertex.alloc(static_cast<std::size_t>(c->n_owned) + c->n_ghost);
  const double hx = 1.0 / static_cast<double>(nx), hy = 1.0 / static_cast<double>(ny),
               hz = 1.0 / static_cast<double>(nz);
  box_vertices<<<static_cast<unsigned>((n_vertices + BX_THREADS - 1) / BX_THREADS), BX_THREADS, 0, c->stream>>>(
      B, hx, hy, hz, c->xyz3.p, c->xyz.p, c->dof_vertex.p);
  box_cells_p1<<<static_cast<unsigned>((n_cubes + BX_THREADS - 1) / BX_THREADS), BX_THREADS, 0, c->stream>>>(
      B, c->x_dofmap.p, c->dofmap.p);
  PTB_CUDA(cudaGetLastError());
  c->launches += 2;
}
#endif // PTB_HOST_EMU

} // namespace ptb
