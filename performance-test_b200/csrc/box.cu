// Device-side mesh and Lagrange P1-P3 function space of the reference's unit cube -- SURVEY 8f row 4: the
// "ZZZ Create Mesh" / "ZZZ FunctionSpace" regions (mesh.cpp:184-186 mesh::create_box(tetrahedron),
// poisson_problem.cpp:33-47 create_functionspace) generated on the device instead of being passed
// in through ptb_set_mesh / ptb_set_space, so that a C5-scale slab (198 M cells per GPU) needs no
// multi-GB host staging. Same arrays as the host stand-in (host/box_mesh.cpp, host/fem.cpp), bit
// for bit: Kuhn 6-tet split of every cube (all six share the body diagonal 0-7), vertices
// lexicographic (plane, iy, ix), x = ix * (1 / nx); z-slab partition with one ghost layer of cells
// below the slab; dofs numbered level-major by entity kind (common/kuhn_space.h), owned levels
// first, then the ghost level below, then the ghost plane block above. For P1 the dofs are the
// vertices; for P2/P3 the kernels also produce the dof coordinates (V->tabulate_dof_coordinates()).
// Run on the B200 since round 2 (GPU tests compare every array with the stand-in's bit for bit); tests/emu
// also runs these sources on the host against the stand-in's arrays.
#include "../common/kuhn_space.h"
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
  std::int64_t G0, G1, Glow, Ghigh; // global dof numbers: owned [G0, G1), ghosts [Glow, G0) and [G1, Ghigh)
};

// The numbering of common/kuhn_space.h (Numbering + the per-tet-type local dof table) by value.
struct SpaceDims
{
  int order, nd;
  std::int64_t PS, LS;                 // dofs of a plane block / a layer block; level stride = PS + LS
  std::int64_t koff[kuhn::NK], kw[kuhn::NK];
  std::int32_t ksub[kuhn::NK];
  std::uint8_t kdim[kuhn::NK], kd1[kuhn::NK], kd2[kuhn::NK], klayer[kuhn::NK];
  std::uint8_t tkind[120], tbx[120], tby[120], tbz[120], tsub[120]; // [tet type][local dof]
  double edge_t[2];                    // GLL-warped edge parameters of the order
};

__device__ __forceinline__ std::int64_t space_global(const SpaceDims& N, int k, std::int64_t level,
                                                     std::int64_t iy, std::int64_t ix, int sub)
{
  return level * (N.PS + N.LS) + (N.klayer[k] ? N.PS : 0) + N.koff[k] + (iy * N.kw[k] + ix) * N.ksub[k] + sub;
}

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
__global__ void box_vertices(BoxDims B, std::int64_t level_stride, double hx, double hy, double hz,
                             double* __restrict__ xyz3, double* __restrict__ xyz4,
                             std::int32_t* __restrict__ dof_vertex)
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
  // vertex dofs lead the plane block of their level (kind 0, one dof each)
  const std::int32_t d = box_to_local(B, pz * level_stride + iy * nvx + ix);
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

// thread per local tetrahedron, any order: dof of local index i = the entity (kind, base, sub) the
// per-tet-type table names, numbered by space_global, mapped into the local ranges
__global__ void box_cells_dofmap(BoxDims B, SpaceDims N, std::int32_t* __restrict__ dofmap, int* __restrict__ flags)
{
  const std::int64_t cell = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (cell >= 6 * B.nx * B.ny * (B.l1 - B.l0))
    return;
  const int t = static_cast<int>(cell % 6);
  const std::int64_t cube = cell / 6;
  const std::int64_t ix = cube % B.nx, iy = (cube / B.nx) % B.ny, iz = cube / (B.nx * B.ny) + B.l0;
  for (int i = 0; i < N.nd; ++i)
  {
    const int e = t * N.nd + i;
    const std::int32_t l = box_to_local(
        B, space_global(N, N.tkind[e], iz + N.tbz[e], iy + N.tby[e], ix + N.tbx[e], N.tsub[e]));
    if (l < 0)
      flags[0] = 1; // cell dof outside the local ranges
    dofmap[cell * N.nd + i] = l;
  }
}

// thread per local dof: invert the numbering (level, block, kind, base, sub) and place the dof on
// its entity -- vertices at the lattice point, edge dofs at the GLL-warped parameter from the lower
// vertex, face dofs at the centroid (host/fem.cpp create_functionspace, same expressions)
__global__ void box_dof_coordinates(BoxDims B, SpaceDims N, double hx, double hy, double hz, double* __restrict__ dof_x)
{
  const std::int64_t n_owned = B.G1 - B.G0, n_low = B.G0 - B.Glow, n = n_owned + n_low + (B.Ghigh - B.G1);
  const std::int64_t l = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (l >= n)
    return;
  const std::int64_t g = l < n_owned ? B.G0 + l : (l < n_owned + n_low ? B.Glow + (l - n_owned) : B.G1 + (l - n_owned - n_low));
  const std::int64_t S = N.PS + N.LS, level = g / S;
  std::int64_t rem = g % S;
  const bool layer = rem >= N.PS;
  if (layer)
    rem -= N.PS;
  int k = -1;
  for (int q = 0; q < kuhn::NK; ++q)
    if (N.ksub[q] > 0 && (N.klayer[q] != 0) == layer && rem >= N.koff[q])
      k = q; // kinds of a block are laid out in ascending q: the last one that starts at or below rem
  const std::int64_t idx = (rem - N.koff[k]) / N.ksub[k];
  const int sub = static_cast<int>((rem - N.koff[k]) % N.ksub[k]);
  const std::int64_t iy = idx / N.kw[k], ix = idx % N.kw[k];
  double ox = 0, oy = 0, oz = 0; // offset from the base in lattice units
  const int d1 = N.kd1[k], d2 = N.kd2[k];
  if (N.kdim[k] == 1)
  {
    const double t = N.edge_t[sub];
    ox = __dmul_rn(t, static_cast<double>(d1 & 1)), oy = __dmul_rn(t, static_cast<double>((d1 >> 1) & 1)),
    oz = __dmul_rn(t, static_cast<double>((d1 >> 2) & 1));
  }
  else if (N.kdim[k] == 2)
  {
    ox = ((d1 & 1) + (d2 & 1)) / 3.0;
    oy = (((d1 >> 1) & 1) + ((d2 >> 1) & 1)) / 3.0;
    oz = (((d1 >> 2) & 1) + ((d2 >> 2) & 1)) / 3.0;
  }
  dof_x[3 * l + 0] = __dmul_rn(hx, __dadd_rn(static_cast<double>(ix), ox));
  dof_x[3 * l + 1] = __dmul_rn(hy, __dadd_rn(static_cast<double>(iy), oy));
  dof_x[3 * l + 2] = __dmul_rn(hz, __dadd_rn(static_cast<double>(level), oz));
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

// The numbering of the order by value (common/kuhn_space.h: Numbering + local table).
SpaceDims make_space_dims(std::int64_t nx, std::int64_t ny, std::int64_t nz, int order)
{
  const kuhn::Numbering K(nx, ny, nz, order);
  std::vector<kuhn::LocalDof> tab;
  kuhn::build_local_table(order, tab);
  SpaceDims N{};
  N.order = order, N.nd = kuhn::lagrange_ndofs(order);
  N.PS = K.PS, N.LS = K.LS;
  for (int k = 0; k < kuhn::NK; ++k)
  {
    N.koff[k] = K.koff[k], N.kw[k] = K.kw[k], N.ksub[k] = K.ksub[k];
    N.kdim[k] = static_cast<std::uint8_t>(kuhn::kinds[k].dim), N.kd1[k] = static_cast<std::uint8_t>(kuhn::kinds[k].d1);
    N.kd2[k] = static_cast<std::uint8_t>(kuhn::kinds[k].d2), N.klayer[k] = kuhn::kinds[k].layer ? 1 : 0;
  }
  for (std::size_t e = 0; e < tab.size(); ++e)
  {
    N.tkind[e] = static_cast<std::uint8_t>(tab[e].kind), N.tsub[e] = static_cast<std::uint8_t>(tab[e].sub);
    N.tbx[e] = static_cast<std::uint8_t>(tab[e].bx), N.tby[e] = static_cast<std::uint8_t>(tab[e].by);
    N.tbz[e] = static_cast<std::uint8_t>(tab[e].bz);
  }
  for (int s = 0; s < order - 1 && s < 2; ++s)
    N.edge_t[s] = kuhn::edge_param(order, s);
  return N;
}

// The slab of rank `rank` of `nranks`: z-slabs of cube layers, as evenly as possible, one ghost layer
// of cells below (host/box_mesh.cpp slab_range, create_box_mesh; host/fem.cpp local_ranges).
BoxDims make_box_dims(std::int64_t nx, std::int64_t ny, std::int64_t nz, int rank, int nranks, const SpaceDims& N)
{
  const std::int64_t base = nz / nranks, rem = nz % nranks;
  const std::int64_t L0 = rank * base + (rank < rem ? rank : rem), L1 = L0 + base + (rank < rem ? 1 : 0);
  const bool last = rank == nranks - 1;
  const std::int64_t S = N.PS + N.LS;
  BoxDims B{};
  B.nx = nx, B.ny = ny, B.nz = nz;
  B.l0 = rank > 0 ? L0 - 1 : L0, B.l1 = L1;
  B.G0 = L0 * S;
  B.G1 = last ? nz * S + N.PS : L1 * S;
  B.Glow = B.l0 * S;
  B.Ghigh = last ? B.G1 : B.G1 + N.PS;
  return B;
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

// Generates the local slab on the device and fills what ptb_set_mesh and the mesh-dependent half
// of ptb_set_space fill: n_vertices, n_cells, x_dofmap, xyz3, xyz, n_owned, n_ghost, dofmap,
// dof_vertex, and for order > 1 the dof coordinates c->dof_x. Two launches (four for order > 1).
void gpu_create_box(ptb_ctx* c, int order, std::int64_t nx, std::int64_t ny, std::int64_t nz, int rank, int nranks)
{
  const SpaceDims N = make_space_dims(nx, ny, nz, order);
  const BoxDims B = make_box_dims(nx, ny, nz, rank, nranks, N);
  const std::int64_t nvp = (nx + 1) * (ny + 1);
  const std::int64_t n_vertices = nvp * (B.l1 - B.l0 + 1), n_cubes = nx * ny * (B.l1 - B.l0);
  const std::int64_t n_local = B.Ghigh - B.Glow;
  if (n_vertices > INT32_MAX || n_local > INT32_MAX || n_cubes * 6 * N.nd > static_cast<std::int64_t>(UINT32_MAX))
    throw std::runtime_error("ptb_create_box: local slab exceeds 32-bit local indexing");
  c->n_vertices = n_vertices, c->n_cells = 6 * n_cubes;
  c->n_owned = static_cast<std::int32_t>(B.G1 - B.G0);
  c->n_ghost = static_cast<std::int32_t>((B.G0 - B.Glow) + (B.Ghigh - B.G1));
  c->xyz3.alloc(static_cast<std::size_t>(n_vertices) * 3);
  c->xyz.alloc(static_cast<std::size_t>(n_vertices) * 4);
  c->x_dofmap.alloc(static_cast<std::size_t>(n_cubes) * 24);
  c->dofmap.alloc(static_cast<std::size_t>(n_cubes) * 6 * N.nd);
  c->dof_vertex.alloc(static_cast<std::size_t>(n_local));
  PTB_CUDA(cudaMemsetAsync(c->dof_vertex.p, 0xFF, c->dof_vertex.bytes(), c->stream)); // -1: not a vertex dof
  const double hx = 1.0 / static_cast<double>(nx), hy = 1.0 / static_cast<double>(ny),
               hz = 1.0 / static_cast<double>(nz);
  box_vertices<<<static_cast<unsigned>((n_vertices + BX_THREADS - 1) / BX_THREADS), BX_THREADS, 0, c->stream>>>(
      B, N.PS + N.LS, hx, hy, hz, c->xyz3.p, c->xyz.p, c->dof_vertex.p);
  if (order == 1)
  {
    box_cells_p1<<<static_cast<unsigned>((n_cubes + BX_THREADS - 1) / BX_THREADS), BX_THREADS, 0, c->stream>>>(
        B, c->x_dofmap.p, c->dofmap.p);
    c->dof_x.release();
    c->launches += 2;
  }
  else
  {
    DevBuf<std::int32_t> scratch; // box_cells_p1 also writes the P1 dofmap: not wanted here
    DevBuf<int> flags;
    scratch.alloc(static_cast<std::size_t>(n_cubes) * 24);
    flags.alloc(1);
    flags.zero(c->stream);
    box_cells_p1<<<static_cast<unsigned>((n_cubes + BX_THREADS - 1) / BX_THREADS), BX_THREADS, 0, c->stream>>>(
        B, c->x_dofmap.p, scratch.p);
    box_cells_dofmap<<<static_cast<unsigned>((6 * n_cubes + BX_THREADS - 1) / BX_THREADS), BX_THREADS, 0, c->stream>>>(
        B, N, c->dofmap.p, flags.p);
    c->dof_x.alloc(static_cast<std::size_t>(n_local) * 3);
    box_dof_coordinates<<<static_cast<unsigned>((n_local + BX_THREADS - 1) / BX_THREADS), BX_THREADS, 0, c->stream>>>(
        B, N, hx, hy, hz, c->dof_x.p);
    int bad = 0;
    PTB_CUDA(cudaMemcpyAsync(&bad, flags.p, sizeof(bad), cudaMemcpyDeviceToHost, c->stream));
    PTB_CUDA(cudaStreamSynchronize(c->stream)); // scratch, flags die here
    if (bad)
      throw std::runtime_error("ptb_create_box: cell dof outside the local ranges");
    c->launches += 4;
  }
  PTB_CUDA(cudaGetLastError());
}
#endif // PTB_HOST_EMU

} // namespace ptb
